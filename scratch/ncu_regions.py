"""Instruction / stall-sample breakdown of one kernel by named source-line regions.
usage: ncu_regions.py <report.ncu-rep> <lib.so> <mangled-func-substring> <source-file-substring> <pixels> <regions.txt>
regions.txt lines: lo hi name   (line ranges of the source file)"""
import csv, re, sys, collections, subprocess, os, tempfile
rep, lib, func, srcfile, px, regfile = sys.argv[1:7]
px = int(px)
tmp = tempfile.mkdtemp()
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(tmp, "src.csv"), "w").write(raw)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if func in out:
            dis = out
            break
regions = []
for ln in open(regfile):
    a = ln.split(None, 2)
    if len(a) == 3:
        regions.append((int(a[0]), int(a[1]), a[2].strip()))
cur = None; infunc = False; addr2line = {}
for ln in dis.splitlines():
    if ln.startswith('.text.') and ln.strip().endswith(':'):
        infunc = func in ln; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ia = hdr.index('Address'); ii = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
base = None; inst = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0; ops = collections.Counter()
lines = collections.Counter(); lsamp = collections.Counter()
for r in rows[2:]:
    try: a = int(r[ia], 16); n = int(r[ii]); s = int(r[isamp])
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base)
    name = 'other'
    if key and srcfile in key[0]:
        for lo, hi, nm in regions:
            if lo <= key[1] <= hi: name = nm
        lines[key[1]] += n; lsamp[key[1]] += s
    elif key: name = key[0]
    inst[name] += n; samp[name] += s; tot += n; tots += s
    t = r[isrc].split()
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += n
print(f"total warp-instr {tot}  = {tot*32/px:.1f} thread-instr/px; samples {tots}")
for k, n in inst.most_common():
    print(f"{k:44s} {n/tot*100:6.2f}% instr {n*32/px:7.1f} /px  {samp[k]/max(1,tots)*100:6.2f}% samples")
print()
print("  ".join(f"{k}:{n*32/px:.1f}" for k, n in ops.most_common(28)))
print()
for l, n in lines.most_common(25):
    print(f"line {l:5d}  {n*32/px:6.1f}/px  {lsamp[l]/max(1,tots)*100:5.2f}% samples")
