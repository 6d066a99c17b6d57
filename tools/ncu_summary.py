"""Summarise ncu outputs into small text/JSON files for profiles/ (run here, no GPU needed).
usage: python tools/ncu_summary.py <launches.csv> <prof.ncu-rep> <out_prefix>"""
import csv, json, statistics, subprocess, sys
from collections import defaultdict

launches, rep, out = sys.argv[1:4]
rows = [l for l in open(launches) if l.startswith('"')]
g = defaultdict(list)
for r in csv.DictReader(rows):
    try:
        g[(r["Kernel Name"].split("(")[0], r["Grid Size"], r["Block Size"])].append(float(r["Metric Value"].replace(",", "")))
    except Exception:
        pass
tot = sum(sum(v) for v in g.values())
lines = ["kernel,grid,block,launches,total_us,share,mean_us,median_us,min_us,max_us"]
for (k, gs, bs), v in sorted(g.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k},{gs.replace(',', ' ')},{bs.replace(',', ' ')},{len(v)},{sum(v)/1e3:.1f},{sum(v)/tot:.4f},{statistics.mean(v)/1e3:.2f},"
                 f"{statistics.median(v)/1e3:.2f},{min(v)/1e3:.2f},{max(v)/1e3:.2f}")
open(out + "_launches.csv", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max"]
summ = []
for d in data:
    e = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            e[w] = f"{d[i]} {units[i]}".strip()
    summ.append(e)
json.dump(summ, open(out + "_full.json", "w"), indent=1)
print(json.dumps(summ[:1], indent=1))
