"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mpopis::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] in ("ns", "nsecond") else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print(f"| kernel | launches | total µs | µs/launch | share |\n|---|---:|---:|---:|---:|")
for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {v:.1f} | {v / n:.2f} | {100 * v / tot:.1f}% |")
print(f"| **total** | {sum(n for n, _ in agg.values())} | {tot:.1f} | | |")
