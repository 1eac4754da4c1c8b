"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line.
usage: python scratch/ncu_lines.py dump.csv [min_pct]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
cur_file = None
agg = collections.OrderedDict(); last = ("?", "0"); agg[last] = [0.0, 0.0, ""]
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or len(r) <= iI: continue
    line, src = r[0], r[1]
    key = (cur_file, line)
    try:
        v = float(r[iI] or 0); w = float(r[iW] or 0)
    except ValueError:
        continue
    if line:
        last = key
        if key not in agg: agg[key] = [0.0, 0.0, src]
    else:
        key = last
    agg[key][0] += v; agg[key][1] += w
tot = sum(a[0] for a in agg.values()); totw = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot:.0f}, stall samples {totw:.0f}")
for (f, l), (v, w, s) in agg.items():
    if v > thr / 100 * tot or w > thr / 100 * totw:
        print(f"{f:12s}:{l:>4} inst {v/tot*100:5.1f}%  stall {w/totw*100:5.1f}%  {s.strip()[:100]}")
