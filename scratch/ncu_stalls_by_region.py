"""Stall-reason samples per source region. usage: ncu_stalls_by_region.py src.csv lib.so func srcfile regions.txt"""
import csv, re, sys, collections, subprocess, os, tempfile
srccsv, lib, func, srcfile, regfile = sys.argv[1:6]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if func in out: dis = out; break
regions = [(int(a[0]), int(a[1]), a[2].strip()) for a in (ln.split(None, 2) for ln in open(regfile)) if len(a) == 3]
cur = None; infunc = False; addr2line = {}
for ln in dis.splitlines():
    if ln.startswith('.text.') and ln.strip().endswith(':'):
        infunc = func in ln; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
ia = hdr.index('Address')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None
tab = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base)
    name = 'other'
    if key and srcfile in key[0]:
        for lo, hi, nm in regions:
            if lo <= key[1] <= hi: name = nm
    elif key: name = key[0]
    for i, h in stall_cols:
        try: tab[name][h] += int(r[i])
        except Exception: pass
tot = sum(sum(c.values()) for c in tab.values())
cols = ['stall_selected','stall_wait','stall_short_sb','stall_long_sb','stall_mio','stall_barrier','stall_math','stall_lg','stall_not_selected','stall_branch_resolving','stall_dispatch','stall_no_inst']
print(f"{'region':28s} {'tot%':>6s} " + " ".join(f"{c[6:12]:>7s}" for c in cols))
for name, c in sorted(tab.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(c.values())
    print(f"{name:28s} {s/tot*100:6.2f} " + " ".join(f"{c[k]/tot*100:7.2f}" for k in cols))
allc = collections.Counter()
for c in tab.values(): allc.update(c)
print(f"{'ALL':28s} {100:6.2f} " + " ".join(f"{allc[k]/tot*100:7.2f}" for k in cols))
