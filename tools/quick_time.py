"""Quick device timing of one config: python tools/quick_time.py c3 5"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kube_batch_b200 import engine, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
snap, conf = synth.make(name)
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = engine.Engine(0, flags=flags)
t0 = time.time(); eng.load(snap, conf); t1 = time.time()
print(f"{name}: T={snap.T} N={snap.N} J={snap.J} load {1e3*(t1-t0):.1f} ms")
for i in range(reps):
    t0 = time.time(); r = eng.allocate(); t1 = time.time()
    s = r.stats
    print(f"  rep{i}: wall {1e3*(t1-t0):.2f} ms gpu {s.gpu_ms:.2f} ms launches {s.kernel_launches} visits {s.visits} "
          f"scans {s.pairs_scanned // max(1, snap.N)} classes {s.n_classes} processed {s.tasks_processed} alloc {s.tasks_allocated} "
          f"pipe {s.tasks_pipelined} ready {s.jobs_ready} replayed {s.pairs_replayed} "
          f"logical pairs/s {s.pairs_logical / (s.gpu_ms * 1e-3):.3e}\n"
          f"        predictions {s.predictions} mispredictions {s.mispredictions}\n"
          f"        scans {s.scans} rescans {s.rescans} per-launch cycles: scan {s.cyc_scan / max(1, s.scans):.0f} merge {s.cyc_merge / max(1, s.scans):.0f} "
          f"replay+ctl {s.cyc_replay / max(1, s.scans):.0f} (steps {s.cyc_steps / max(1, s.scans):.0f} ctl {s.cyc_ctl / max(1, s.scans):.0f}) total {s.cyc_total / max(1, s.scans):.0f}; us/launch {1e3 * s.gpu_ms / max(1, s.kernel_launches):.2f}\n"
          f"        timing mode (KB_PIPE_TIMING=1): step loop {s.cyc_merge / max(1, s.scans):.0f} hot ring {s.cyc_ring / max(1, s.scans):.0f} plan command {s.cyc_plan / max(1, s.scans):.0f} "
          f"runs {s.predictions}; list wait {s.cyc_total / max(1, s.scans):.0f} cycles per visit chain, {s.mispredictions} visits waited\n"
          f"        pipeline {s.pipeline}: requests {s.pipe_requests} urgent {s.pipe_urgent} extends {s.pipe_extends} patched lists {s.pipe_patched} "
          f"patch entries {s.pipe_patch_entries}; us/visit-chain {1e3 * s.gpu_ms / max(1, s.scans):.2f}")
