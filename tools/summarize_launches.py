"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or not r[0].isdigit():
        continue
    name = r[ki].split("(")[0].replace("dynmm::<unnamed>::", "")
    name = name[:60]
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':62s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:62s} {n:8d} {t / 1e3:10.1f} {100 * t / tot:6.1f}%")
print(f"{'TOTAL':62s} {sum(a[0] for a in agg.values()):8d} {tot / 1e3:10.1f}")
