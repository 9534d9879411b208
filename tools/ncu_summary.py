import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:48]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in d.items(): print(k, len(v), "avg_us %.1f"%(sum(v)/len(v)/1e3), "min %.1f max %.1f"%(min(v)/1e3, max(v)/1e3))
