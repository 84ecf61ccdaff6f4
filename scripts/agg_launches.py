"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]; ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
    k = r[ki][:100]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"total {tot/div:.1f} us per step over {sum(n for n,_ in agg.values())/div:.0f} launches/step")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print(f"{v/div:9.1f} us  x{n/div:5.1f}  {v/tot*100:5.1f}%  avg {v/n:7.1f} us  {k}")
