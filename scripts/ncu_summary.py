"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit in ("ns", "nsecond"): v /= 1e3
    elif unit in ("ms", "msecond"): v *= 1e3
    elif unit in ("s", "second"): v *= 1e6
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
total = sum(v[1] for v in tot.values())
print(f"| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100*t/total:.1f}% | {t/n:.2f} |")
print(f"\ntotal {total:.1f} us over {sum(v[0] for v in tot.values())} launches")
