"""ncu source page (csv) x nvdisasm -g: instructions, samples and stall reasons per named source region.
usage: ncu_regions2.py src.csv all.dis kernel_name file:lo-hi=name ..."""
import csv, re, sys, collections
sass_csv, disasm, func = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for a in sys.argv[4:]:
    rng, name = a.split("=")
    f, lh = rng.split(":")
    lo, hi = lh.split("-")
    regions.append((f, int(lo), int(hi), name))
cur = None; infunc = False; addr2line = {}
for ln in open(disasm):
    if ln.startswith('.text.') and ln.strip().endswith(':'):
        infunc = func in ln
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia = hdr.index('Address'); ii = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def region_of(key):
    if key is None: return "?"
    for f, lo, hi, name in regions:
        if key[0].endswith(f) and lo <= key[1] <= hi: return name
    return f"{key[0]}:other"
base = None
inst = collections.Counter(); samp = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
tot = tots = 0
for r in rows[2:]:
    try: a = int(r[ia], 16); n = int(r[ii]); s = int(r[isamp])
    except Exception: continue
    if base is None: base = a
    reg = region_of(addr2line.get(a - base))
    inst[reg] += n; samp[reg] += s; tot += n; tots += s
    for i, h in stall_cols:
        try: stalls[reg][h] += int(r[i])
        except Exception: pass
print(f"total warp instructions {tot}, samples {tots}")
allst = collections.Counter()
for reg, n in inst.most_common():
    top = ", ".join(f"{h[6:]} {v / max(1, samp[reg]) * 100:.0f}%" for h, v in stalls[reg].most_common(4))
    print(f"{reg:28s} {n / tot * 100:6.2f}% instr {samp[reg] / tots * 100:6.2f}% samples | {top}")
    allst.update(stalls[reg])
print("all:", ", ".join(f"{h[6:]} {v / tots * 100:.1f}%" for h, v in allst.most_common(10)))
