"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u in ("ms", "msecond") else 1.0)   # -> us
    name = r[ki][:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{sum(a[0] for a in agg.values())} launches, {tot/1e3:.2f} ms of kernel time\n")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {t/1e3:.2f} | {t/n:.1f} | {100*t/tot:.1f} % |")
