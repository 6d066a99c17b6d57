"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    agg[row["Kernel Name"][:70]].append(v)
tot = sum(sum(v) for v in agg.values())
print("kernel,launches,avg_us,median_us,max_us,sum_ms,share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    s = sorted(v)
    print(f"\"{k}\",{len(v)},{sum(v) / len(v):.2f},{s[len(v) // 2]:.2f},{s[-1]:.2f},{sum(v) / 1000:.3f},{sum(v) / tot:.4f}")
