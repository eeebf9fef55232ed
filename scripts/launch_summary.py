"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python scripts/launch_summary.py launches.csv [skip_first_n_launches_per_kernel]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) < 10: continue
    d[r[4].split('(')[0][:44]].append(float(r[-1].replace(',', '')) / 1000)
d = {k: [max(v)] * len(v) for k, v in d.items()}  # the full-batch launch (chunked e2e launches are smaller)
tot = sum(v[-1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][-1]):
    print(f"{k:46s} n={len(v):3d} max={v[-1]:8.1f} us  {100*v[-1]/tot:5.1f}%")
print("sum of last launches", round(tot, 1), "us")
