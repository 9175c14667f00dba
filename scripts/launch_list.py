"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, ns, share)."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'launches':>8s} {'ns total':>14s} {'share':>7s}")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {n:8d} {int(ns):14d} {100 * ns / tot:6.2f}%")
print(f"{'total':70s} {len(rows):8d} {int(tot):14d}")
