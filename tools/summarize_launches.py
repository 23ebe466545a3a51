"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    name = name.split('(')[0][-70:]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}.get(unit, v)
    a = agg[name]; a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(v[1] for v in agg.values())
print(f"{'total ms':>10} {'share':>6} {'count':>6} {'avg us':>9} {'max us':>9}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3:10.2f} {100*v[1]/tot:5.1f}% {v[0]:6d} {v[1]/v[0]:9.1f} {v[2]:9.1f}  {k}")
print(f"{tot/1e3:10.2f} ms total over {sum(v[0] for v in agg.values())} launches")
