"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST pass (second half)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = rows[hdr + 1:]
data = data[len(data) // 2:]
agg = collections.OrderedDict()
for r in data:
    name = r[kn].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    t = float(r[mv].replace(",", "")) / 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {tot:.0f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {n:4d} {t:8.1f} us {100 * t / tot:5.1f} %  avg {t / n:6.1f}")
