"""One allocate cycle of a seeded session with inter-pod (anti)affinity terms on cuda:0 (run under ncu for the launch list:
visit_kernel<0,1> + aff_prepass_kernel<0..2>).  python tools/affinity_run.py [nodes] [groups] [repeats]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aff_gen  # noqa: E402  (the seeded object-level generator; no oracle, no emulation involved)
from kube_batch_b200 import engine  # noqa: E402
from kube_batch_b200.snapshot import PluginConf  # noqa: E402

if __name__ == "__main__":
    nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    snap = aff_gen.random_affinity_session(4242, n_nodes=nodes, n_groups=groups, p_affine=0.35, spec_pool=6).flatten()
    eng = engine.Engine(device=0)
    t0 = time.perf_counter()
    eng.load(snap, PluginConf.default())
    load_ms = 1e3 * (time.perf_counter() - t0)
    out = []
    for _ in range(reps):
        r = eng.allocate()
        st = r.stats
        out.append({"gpu_ms": float(st.gpu_ms), "scans": int(st.scans), "kernel_launches": int(st.kernel_launches), "allocated": int(st.tasks_allocated),
                    "pipelined": int(st.tasks_pipelined), "pairs_logical": int(st.pairs_logical), "pairs_scanned": int(st.pairs_scanned)})
    print(json.dumps({"nodes": int(snap.N), "tasks": int(snap.T), "counter_groups": int(snap.pod_affinity["n_groups"]),
                      "pod_kinds": int(snap.pod_affinity["n_kinds"]), "load_ms": load_ms, "cycles": out}))
    eng.close()
