"""Print SASS instructions (with executed counts and stall samples) that map to a range of CUDA source lines.
usage: ncu_sass_lines.py <source.csv> <nvdisasm -g output> <func substring> <line_lo> <line_hi>"""
import csv, re, sys
sass_csv, disasm, func, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
cur=None; infunc=False; addr2line={}
for ln in open(disasm):
    if ln.startswith('.text.') and ln.strip().endswith(':'):
        infunc = func in ln; continue
    if not infunc: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(sass_csv)))
hdr=rows[1]; ia=hdr.index('Address'); ii=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); isrc=hdr.index('Source')
base=None
for r in rows[2:]:
    try: a=int(r[ia],16); n=int(r[ii]); s=int(r[isamp])
    except Exception: continue
    if base is None: base=a
    key=addr2line.get(a-base)
    if key and key[0].startswith('fsg_topousm_v6') and lo<=key[1]<=hi:
        print(f"{a-base:06x} L{key[1]:4d} {n:10d} {s:6d}  {r[isrc].strip()}")
