"""Aggregate `ncu --page source --csv` output over many launches: sum samples per SASS address, print the hottest.
usage: ncu -i rep --page source --csv | python tools/ncu_src_agg.py [topN]"""
import csv, sys
from collections import defaultdict
top = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rows = list(csv.reader(sys.stdin))
agg = {}
order = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        hdr = None
        continue
    if hdr is None:
        hdr = r
        ci = {n: i for i, n in enumerate(hdr)}
        stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    try:
        addr = r[ci["Address"]]; n = int(r[ci["# Samples"]]); ex = int(r[ci["Instructions Executed"]])
    except Exception:
        continue
    if addr not in agg:
        agg[addr] = {"src": r[ci["Source"]].strip(), "samples": 0, "exec": 0, "stalls": defaultdict(int), "idx": len(order)}
        order.append(addr)
    a = agg[addr]; a["samples"] += n; a["exec"] += ex
    for s in stalls:
        v = r[ci[s]]
        if v.isdigit() and int(v): a["stalls"][s] += int(v)
tot = sum(a["samples"] for a in agg.values())
print("total samples", tot, "instructions", len(agg))
# region split by executed count relative to max
mx = max(a["exec"] for a in agg.values()) or 1
for name, lo, hi in (("all-warps", 0.5, 2), ("few-warps", 0.02, 0.5), ("one-warp", 0, 0.02)):
    ss = sum(a["samples"] for a in agg.values() if lo * mx <= a["exec"] < hi * mx)
    print(f"region {name}: samples {ss} ({100.0 * ss / max(1, tot):.1f}%)")
for addr in sorted(agg, key=lambda k: -agg[k]["samples"])[:top]:
    a = agg[addr]
    print(a["idx"], a["src"][:70].ljust(70), a["samples"], "exec", a["exec"], dict(a["stalls"]))
