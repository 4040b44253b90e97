"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: count, mean and share."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    n = r[ki].split('(')[0][:70]; v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns':
        v /= 1e3
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
for n, v in agg.items():
    print(f"{n:70s} n={len(v):4d} avg={sum(v)/len(v):9.1f} us  share={100*sum(v)/tot:5.1f}%")
