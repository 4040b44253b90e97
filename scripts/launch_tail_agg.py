import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; H=rows[h]; ki=H.index("Kernel Name"); vi=H.index("Metric Value"); ui=H.index("Metric Unit")
L=[(r[ki].split("(")[0][:60], float(r[vi].replace(",",""))/(1e3 if r[ui]=="ns" else 1)) for r in rows[h+1:] if len(r)>vi]
n=int(sys.argv[2]) if len(sys.argv)>2 else 600
tail=L[-n:-30]
agg=collections.OrderedDict()
for k,v in tail: agg.setdefault(k,[]).append(v)
tot=sum(sum(v) for v in agg.values())
print("launches", len(L), "tail total us", tot)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print("%-62s n=%4d avg=%8.1f us share=%5.1f%%"%(k,len(v),sum(v)/len(v),100*sum(v)/tot))
