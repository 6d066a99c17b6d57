"""Inter-pod (anti)affinity fuzz beyond the pytest suite: seeded object-level sessions (tests/aff_gen.py) -> CPU emulation of the
device algorithm vs the oracle, which walks the raw pod objects (predicates.go:1261-1572, interpod_affinity.go:99-235).
Three generators: arbitrary required / preferred terms (counter path), the same plus preferred NODE affinity, and host-level
anti-affinity only (groups as port-word atoms: every launch mode incl. the pipeline protocol, and the counter path via KB_AFF_ATOMS=0).
Round 2: 2994 + 2000 + 1997 sessions x 4 / 5 / 3 tier configurations x {allocate + backfill}, 0 mismatches.
usage: python tests/fuzz_affinity.py [first_seed [count]]"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
from kube_batch_b200.snapshot import PluginConf  # noqa: E402
from oracle import kbo  # noqa: E402
import aff_gen  # noqa: E402
import util  # noqa: E402
from test_pod_affinity import AFF_CONFS  # noqa: E402

BOTH = AFF_CONFS + [PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "-3", "podaffinity.weight": "2"}})]
HOST = [PluginConf.default(), PluginConf.from_names([["gang"], ["predicates"]]),
        PluginConf.from_names([["priority", "gang"], ["drf", "predicates", "proportion", "nodeorder"]], {"nodeorder": {"podaffinity.weight": "0"}})]


def one(what, snap, confs, modes):
    bad = 0
    for ci, conf in enumerate(confs):
        o = kbo.allocate(snap, conf, actions=3)
        for mode in modes:
            try:
                e = util.emu_allocate(snap, conf, actions=3, mode=mode)
                util.assert_same_decisions(o.decisions, e.decisions, f"{what} conf {ci} mode {mode}")
                st = util.emu_states(e)
                util.assert_same_state(o, st[0], st[1], f"{what} conf {ci} mode {mode}")
            except Exception as ex:          # noqa: BLE001
                bad += 1
                print("FAIL", str(ex)[:300], flush=True)
    return bad


if __name__ == "__main__":
    lo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    bad = n = 0
    for seed in range(lo, lo + cnt):
        s = aff_gen.random_affinity_session(seed, n_nodes=4 + seed % 13, n_groups=3 + seed % 6, besteffort=seed % 4 == 3).flatten()
        if s.pod_affinity is not None:
            bad += one(f"terms seed {seed}", s, AFF_CONFS, (1,)); n += 1
        s = aff_gen.random_affinity_session(seed, n_nodes=4 + seed % 13, n_groups=3 + seed % 6, node_pref=True, p_affine=0.6 if seed % 3 else 0.0).flatten()
        bad += one(f"terms + node preference seed {seed}", s, BOTH, (1,)); n += 1
        pg = seed % 2 == 0
        s = aff_gen.host_spread_session(seed, n_nodes=3 + seed % 14, n_groups=3 + seed % 9, pipe_geometry=pg, ports=seed % 3 == 0).flatten(W=2 if pg else 1)
        if s.pod_affinity is not None:
            bad += one(f"host spread seed {seed}", s, HOST, (0, 1, 5)); n += 1
            os.environ["KB_AFF_ATOMS"] = "0"
            bad += one(f"host spread (counter path) seed {seed}", s, HOST, (1,))
            os.environ.pop("KB_AFF_ATOMS", None)
    print(f"{n} sessions, {bad} mismatches")
    sys.exit(1 if bad else 0)
