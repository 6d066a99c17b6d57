"""The N>1 path on CPU: world_size 2 and 3 over torch.distributed/gloo.  Each rank scans only its node
shard, the ranks all-gather their top-32 candidate keys together with the candidates' node records, and
every rank replays identically — results must equal the single-process oracle on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kube_batch_b200 import synth  # noqa: E402
from kube_batch_b200.snapshot import PluginConf  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    import util
    from oracle import kbo
    snap, conf = case()
    actions = getattr(case, "actions", 1)
    out = util.emu_sharded_rank(rank, world, port, snap, conf, actions=actions)
    ref = kbo.allocate(snap, conf, actions=actions)
    try:
        util.assert_same_decisions(ref.decisions, out.decisions, f"rank{rank}/{world}")
        ns, os_ = util.emu_states(out)
        util.assert_same_state(ref, ns, os_, f"rank{rank}/{world}")
        q.put((rank, "ok", int(out.result.scans)))
    except AssertionError as e:
        q.put((rank, "FAIL: " + str(e)[:500], 0))


def case_c2():
    return synth.make("c2")


def case_multi_tile_multi_queue():
    s = synth.generate(synth.SynthSpec("mr", tasks=700, jobs=70, nodes=1100, queues=3, hetero_job_frac=0.3, prio_levels=2,
                                       min_member_frac=0.5, seed=4242))
    return s, PluginConf.default()


def case_fewer_tiles_than_ranks():
    return synth.random_session(5, tasks=80, jobs=9, nodes=40, queues=2), PluginConf.default()


def case_allocate_then_backfill():
    s = synth.random_session(11, tasks=400, jobs=30, nodes=700, queues=2, min_member_frac=0.5, be_frac=0.3, be_variants=True)
    return s, PluginConf.default()


case_allocate_then_backfill.actions = 3      # "allocate, backfill": kb_allocate then kb_backfill on every rank


@pytest.mark.parametrize("world,case", [(2, case_c2), (2, case_allocate_then_backfill), (2, case_multi_tile_multi_queue), (3, case_multi_tile_multi_queue),
                                        (2, case_fewer_tiles_than_ranks)])
def test_sharded_node_axis_matches_oracle(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, scans in sorted(res):
        assert status == "ok", (rank, status)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_node_axis_in_process(world):
    """Same protocol with all ranks emulated in one process (the all-gather is a concatenation): cheap enough for many sessions,
    including more ranks than tiles and the allocate + backfill pair."""
    import util
    from oracle import kbo
    from test_emu_parity import CONFS
    for seed in range(10):
        rng = np.random.default_rng(9000 + seed)
        s = synth.random_session(800 + seed, tasks=int(rng.integers(5, 300)), jobs=int(rng.integers(1, 30)), nodes=int(rng.integers(1, 700)),
                                 queues=int(rng.integers(1, 4)), min_member_frac=float(rng.choice([0, .5, 1])), hetero=float(rng.choice([0, .3, 1])))
        cname = list(CONFS)[seed % len(CONFS)]
        for actions in (1, 3):
            o = kbo.allocate(s, CONFS[cname], actions=actions)
            for r, e in enumerate(util.emu_sharded_inprocess(s, CONFS[cname], world, actions=actions)):
                util.assert_same_decisions(o.decisions, e.decisions, f"seed{seed}/{cname}/a{actions}/world{world}/rank{r}")
                ns, os_ = util.emu_states(e)
                util.assert_same_state(o, ns, os_, f"seed{seed}/{cname}/a{actions}/world{world}/rank{r}")


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_preferred_node_affinity_prototype(world):
    """a12 on a sharded node axis (prototype, emulation only): every rank runs pass 1 (max count over the feasible nodes) on its
    replicated copy of the table, so no second exchange is needed; pass 2 (keys) stays sharded."""
    import util
    from oracle import kbo
    from test_emu_parity import _pref_cluster
    conf = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "5"}})
    for seed in range(12):
        s = _pref_cluster(7000 + seed)
        o = kbo.allocate(s, conf)
        for r, e in enumerate(util.emu_sharded_inprocess(s, conf, world, mode=1)):
            util.assert_same_decisions(o.decisions, e.decisions, f"pref seed{seed}/world{world}/rank{r}")
