"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name: count, total, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[start]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 2:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    n = r[ki].split('(')[0].replace('mc::<unnamed>::', '').replace('void mc::', '').replace('void ', '')
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:70]:70s} {v[0]:5d} {v[1] / 1e6:10.3f} ms {100 * v[1] / tot:5.1f}%')
print(f'total {tot / 1e6:.3f} ms')
