"""Summarise an ncu gpu__time_duration launch-list CSV by kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:44]
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
for k, (c, t, m) in agg.items():
    print(f"{k:44s} n={c:5d} total={t/1e6:9.3f} ms avg={t/c/1e3:9.1f} us max={m/1e3:9.1f} us share={100*t/tot:5.1f}%")
