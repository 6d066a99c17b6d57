"""Stall samples of cycle_kernel's MAIN warp by source line (ncu --set full --import-source on).
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_main_warp_lines.py src.csv kb_pipe.cuh-at-capture visits
The main warp is the only executor of the lines between '// ---------------- main warp' and '// ---------------- wind down'."""
import csv, sys
from collections import defaultdict

src_csv, pipe_src, visits = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(src_csv)))
agg = defaultdict(lambda: defaultdict(int)); text = {}
cur = hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ci = {n: i for i, n in enumerate(hdr)}; continue
    if hdr is None or cur is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    key = (cur, ln); text[key] = r[1].strip()[:120]
    for col, name in (("# Samples", "samples"), ("Instructions Executed", "inst")):
        v = r[ci[col]]
        agg[key][name] += int(v) if v.isdigit() else 0
    for s in hdr:
        if s.startswith("stall_") and "Not Issued" not in s and r[ci[s]].isdigit():
            agg[key][s] += int(r[ci[s]])
lines = open(pipe_src).read().split('\n')
lo = next(i for i, l in enumerate(lines) if '// ---------------- main warp ----------------' in l) + 1
hi = next(i for i, l in enumerate(lines) if '// ---------------- wind down ----------------' in l) + 1
main = [(k, v) for k, v in agg.items() if k[0] == 'kb_pipe.cuh' and lo <= k[1] <= hi]
ctl = [(k, v) for k, v in agg.items() if k[0] == 'kb_ctl.h']
mt = sum(v["samples"] for _, v in main); mi = sum(v["inst"] for _, v in main); ci_ = sum(v["inst"] for _, v in ctl)
print(f"# main-warp region kb_pipe.cuh:{lo}-{hi}: {mt} stall samples, {mi} warp-instructions = {mi / visits:.0f} per visit chain "
      f"(+ {ci_ / visits:.0f} in kb_ctl.h: the control plane, lane 0)")
tot = defaultdict(int)
for _, v in main:
    for s, c in v.items():
        if s.startswith("stall_"): tot[s] += c
print("# stall reasons over the region:", ", ".join(f"{s[6:]} {100 * c / max(1, mt):.0f}%" for s, c in sorted(tot.items(), key=lambda x: -x[1])[:8]))
print("# line  samples  share  top stalls | source")
for k, v in sorted(main, key=lambda kv: -kv[1]["samples"])[:40]:
    top = sorted(((s[6:], c) for s, c in v.items() if s.startswith("stall_") and c), key=lambda x: -x[1])[:3]
    print(f"{k[1]:5d} {v['samples']:7d} {100 * v['samples'] / max(1, mt):5.1f}%  {top} | {text[k]}")
