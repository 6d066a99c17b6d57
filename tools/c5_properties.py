"""BASELINE config 5 (1M tasks x 100k nodes) on one GPU: too large for the CPU oracle, so the result is checked through
size-independent properties and the timing is recorded.  python tools/c5_properties.py [c5|c4]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kube_batch_b200 import abi, engine, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
t0 = time.time(); snap, conf = synth.make(name); t1 = time.time()
eng = engine.Engine(0)
eng.load(snap, conf); t2 = time.time()
r = eng.allocate(); t3 = time.time()
r2 = eng.allocate()
d = r.decisions
ns, osr = eng.node_state(), eng.order_state()
checks = {}
checks["repeatable"] = bool(all(np.array_equal(d[f], r2.decisions[f]) for f in ("node", "kind", "step", "dispatched", "dispatch_step")))
checks["no_node_overcommitted"] = bool((ns["idle"][0] > -10).all() and (ns["idle"][1] > -10 * 1024 * 1024).all())
tj = snap.job_of_task()
ready = osr["job_ready"] >= snap.job_min_avail
alloc = d["kind"] == abi.KB_KIND_ALLOCATED
checks["gang_dispatch_iff_ready"] = bool((d["dispatched"][alloc] == ready[tj[alloc]]).all() and not d["dispatched"][~alloc].any())
placed = d["step"] != 0xFFFFFFFF
checks["steps_are_a_permutation"] = bool(np.array_equal(np.sort(d["step"][placed]), np.arange(int(placed.sum()))))
# bookkeeping closes: Used grew by exactly the Resreq of the placed tasks, per node and dimension
used_add = np.zeros_like(ns["used"])
for k in range(snap.R):
    np.add.at(used_add[k], d["node"][placed], snap.task_resreq[k][placed])
checks["node_used_closes"] = bool(np.array_equal(ns["used"], snap.node_used + used_add))
pods_add = np.bincount(d["node"][placed], minlength=snap.N)
checks["pod_counts_close"] = bool(np.array_equal(ns["pods"], snap.node_pods + pods_add))
# every placed task satisfied its selector / taints on the chosen node
n = d["node"][placed]
sel_ok = ((snap.node_labels[:, n] & snap.task_sel_req[:, placed]) == snap.task_sel_req[:, placed]).all()
taint_ok = ((snap.node_taints[:, n] & ~snap.task_tol[:, placed]) == 0).all()
checks["selectors_and_taints_respected"] = bool(sel_ok and taint_ok)
st = r.stats
out = {"config": name, "T": snap.T, "N": snap.N, "J": snap.J, "Q": snap.Q, "gen_s": t1 - t0, "load_ms": 1e3 * (t2 - t1), "gpu_ms": st.gpu_ms,
       "wall_ms": 1e3 * (t3 - t2), "scans": st.scans, "rescans": st.rescans, "visits": st.visits, "classes": st.n_classes,
       "tasks_processed": st.tasks_processed, "tasks_allocated": st.tasks_allocated, "tasks_pipelined": st.tasks_pipelined,
       "podgroups_ready": st.jobs_ready, "pairs_logical": st.pairs_logical, "pairs_scanned": st.pairs_scanned,
       "logical_pairs_per_s": st.pairs_logical / (st.gpu_ms * 1e-3), "scanned_GBps_at_128B": st.pairs_scanned * 128 / (st.gpu_ms * 1e-3) / 1e9,
       "cycles_per_launch": {"scan": st.cyc_scan / max(1, st.scans), "merge": st.cyc_merge / max(1, st.scans), "replay": st.cyc_replay / max(1, st.scans)},
       "checks": checks}
print(json.dumps(out))
sys.exit(0 if all(checks.values()) else 1)
