"""Join an ncu SASS source page (csv) with nvdisasm -g line info: instructions/samples per CUDA line."""
import csv, re, sys, collections
sass_csv, disasm, func = sys.argv[1], sys.argv[2], sys.argv[3]
# parse nvdisasm: track current line; instructions appear as '        /*0010*/  OPCODE ...'
cur=None; infunc=False; addr2line={}
for ln in open(disasm):
    if ln.startswith('.text.') and ln.strip().endswith(':'):
        infunc = func in ln
        continue
    if not infunc: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(sass_csv)))
hdr=rows[1]; ia=hdr.index('Address'); ii=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); isrc=hdr.index('Source')
base=None; inst=collections.Counter(); samp=collections.Counter(); tot=0; tots=0
for r in rows[2:]:
    try: a=int(r[ia],16); n=int(r[ii]); s=int(r[isamp])
    except Exception: continue
    if base is None: base=a
    key=addr2line.get(a-base)
    inst[key]+=n; samp[key]+=s; tot+=n; tots+=s
print("total", tot, tots)
for key,n in inst.most_common(40):
    print(f"{str(key):38s} {n/tot*100:6.2f}% instr {samp[key]/tots*100:6.2f}% samples")
