"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (usage: launch_summary.py file.csv)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < 15 or not r[0].isdigit(): continue
    v = float(r[-1]); v = v / 1000 if r[-2] in ('nsecond', 'ns') else v
    agg.setdefault(r[4][:70], []).append(v)
for k, v in agg.items(): print(f"{k:72s} n={len(v):3d} mean={sum(v)/len(v):8.1f} us")
