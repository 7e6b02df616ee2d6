import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    k = r[ki][:70]; v = float(r[vi].replace(",", ""))
    agg.setdefault(k, []).append(v)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v)/1e6:9.3f} ms total {len(v):4d}x avg {sum(v)/len(v)/1e3:9.1f} us  {k}")
