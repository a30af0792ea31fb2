"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = row["Kernel Name"].split("(")[0].split("::")[-1][-48:]
    tot[name][0] += 1
    tot[name][1] += v
S = sum(v[1] for v in tot.values())
print(f"{'kernel':50s} {'n':>4s} {'total us':>10s} {'us/launch':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {v[0]:4d} {v[1]:10.1f} {v[1]/v[0]:10.1f} {100*v[1]/S:6.1f}%")
print(f"{'total':50s} {sum(v[0] for v in tot.values()):4d} {S:10.1f}")
