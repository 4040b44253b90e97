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
if "--timed-step" in sys.argv:
    # shares inside the second (timed) sweep step of `bench.py --steps 1 --warmup 1`: from its assembly kernel up to the
    # single-RHS SpMV bench that follows the timed region
    L = []
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            v = float(r[vi].replace(',', ''))
            L.append((r[ki].split('(')[0][:70], v / 1e3 if r[ui] == 'ns' else v))
    a = [i for i, (n, _) in enumerate(L) if 'assemble_kernel' in n][1]
    step = L[a:]
    k = [i for i, (n, _) in enumerate(step) if 'spmv_stream_kernel<1' in n]
    step = step[:k[0]] if k else step
    agg = collections.OrderedDict()
    for n, v in step:
        agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n:70s} n={len(v):4d} avg={sum(v)/len(v):9.1f} us  share={100*sum(v)/tot:5.1f}%")
