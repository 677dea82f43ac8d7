"""Summarise an ncu per-launch CSV (gpu__time_duration.sum) by kernel: launches, total time, share of the step."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("ug::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    tot[name][0] += 1
    tot[name][1] += v
s = sum(v[1] for v in tot.values())
print(f"total {s / 1e3:.2f} ms in {sum(v[0] for v in tot.values())} launches (cold-cache, serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k[:60]:60s} n={v[0]:5d}  ms={v[1] / 1e3:8.3f}  share={v[1] / s:6.3f}  avg_us={v[1] / v[0]:8.1f}")
