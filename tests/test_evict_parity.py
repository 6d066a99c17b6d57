"""reclaim / preempt (kb_evict.h) against the oracle.

CPU: the product's host/device-shared algorithm run by tests/emu (one thread) on the reference's own action tests
(preempt_test.go, reclaim_test.go), hand cases and random clusters.  GPU (-m gpu): the same sessions through the C ABI
(kb_session_load_running + kb_reclaim / kb_preempt)."""
from __future__ import annotations

import numpy as np
import pytest

from kube_batch_b200 import abi, builder as B
from kube_batch_b200.snapshot import PluginConf, PluginOption
from oracle import kbo
import util
from test_oracle_golden import _evict_session

EVICT_TIERS = {
    "preempt_test": PluginConf([[PluginOption("conformance", enabled_preemptable=True), PluginOption("gang", enabled_preemptable=True)]]),
    "reclaim_test": PluginConf([[PluginOption("conformance", enabled_reclaimable=True), PluginOption("gang", enabled_reclaimable=True)]]),
}


def golden_sessions():
    """(name, action, snapshot, tiers) of the reference's own preempt / reclaim tests."""
    yield ("preempt_test.go one job", "preempt",
           _evict_session([("preemptee1", "n1", "Running", "pg1"), ("preemptee2", "n1", "Running", "pg1"),
                           ("preemptor1", "", "Pending", "pg1"), ("preemptor2", "", "Pending", "pg1")], ("3", "3Gi"), ["q1"], [("pg1", "q1")]),
           EVICT_TIERS["preempt_test"], 1)
    yield ("preempt_test.go two jobs", "preempt",
           _evict_session([("preemptee1", "n1", "Running", "pg1"), ("preemptee2", "n1", "Running", "pg1"),
                           ("preemptor1", "", "Pending", "pg2"), ("preemptor2", "", "Pending", "pg2")], ("2", "2G"), ["q1"],
                          [("pg1", "q1"), ("pg2", "q1")]),
           EVICT_TIERS["preempt_test"], 2)
    yield ("reclaim_test.go", "reclaim",
           _evict_session([("preemptee1", "n1", "Running", "pg1"), ("preemptee2", "n1", "Running", "pg1"), ("preemptee3", "n1", "Running", "pg1"),
                           ("preemptor1", "", "Pending", "pg2")], ("3", "3Gi"), ["q1", "q2"], [("pg1", "q1"), ("pg2", "q2")]),
           EVICT_TIERS["reclaim_test"], 1)


def random_cluster(seed: int, big: bool = False):
    """Running, terminating and Pending pods of several PodGroups in several queues; some system-critical, some with scalars."""
    rng = np.random.default_rng(7000 + seed)
    b = B.SessionBuilder()
    nq = int(rng.integers(1, 4))
    for q in range(nq):
        b.add_queue(B.Queue(f"q{q}", int(rng.integers(1, 4)), creation=int(rng.integers(0, 3))))
    nn = int(rng.integers(2, 9)) if not big else int(rng.integers(150, 400))
    gpu_nodes = rng.random() < 0.4
    for n in range(nn):
        alloc = {"cpu": 8, "memory": 32e9, "pods": int(rng.integers(6, 14))}
        if gpu_nodes:
            alloc["nvidia.com/gpu"] = 4
        b.add_node(B.Node(f"n{n:04d}", alloc))
    cap = {f"n{n:04d}": 8.0 for n in range(nn)}
    k = 0
    for g in range(int(rng.integers(2, 9)) if not big else int(rng.integers(40, 90))):
        ns = "kube-system" if rng.random() < 0.1 else "ns"
        b.add_pod_group(B.PodGroup(ns, f"g{g}", f"q{int(rng.integers(0, nq))}", min_member=int(rng.integers(0, 4)),
                                   priority=int(rng.integers(0, 3)), creation=int(rng.integers(0, 4))))
        cpu = float(rng.choice([0.5, 1, 2, 3]))
        req = {"cpu": cpu, "memory": cpu * 1e9}
        if gpu_nodes and rng.random() < 0.3:
            req["nvidia.com/gpu"] = 1
        for i in range(int(rng.integers(1, 7))):
            state = rng.choice(["Running", "Running", "Pending", "Pending", "Deleting"])
            node = ""
            if state != "Pending":
                free = [n for n, c in cap.items() if c >= cpu]
                if not free:
                    state = "Pending"
                else:
                    node = str(rng.choice(free))
                    cap[node] -= cpu
            b.add_pod(B.Pod(ns, f"g{g}-p{i}", node, "Pending" if state == "Pending" else "Running", dict(req), group=f"g{g}",
                            priority=int(rng.integers(0, 3)), creation=int(rng.integers(0, 5)) if rng.random() < 0.5 else k,
                            deleting=(state == "Deleting")))
            k += 1
    return b.flatten()


def tier_variants():
    d = PluginConf.default()
    yield "default", d
    yield "preempt_test", EVICT_TIERS["preempt_test"]
    yield "reclaim_test", EVICT_TIERS["reclaim_test"]
    t = True
    yield "one_tier_all", PluginConf([[PluginOption("priority", enabled_job_order=t, enabled_task_order=t, enabled_preemptable=t),
                                       PluginOption("gang", enabled_job_order=t, enabled_job_ready=t, enabled_job_pipelined=t, enabled_preemptable=t, enabled_reclaimable=t),
                                       PluginOption("conformance", enabled_preemptable=t, enabled_reclaimable=t),
                                       PluginOption("drf", enabled_job_order=t, enabled_preemptable=t),
                                       PluginOption("predicates", enabled_predicate=t),
                                       PluginOption("proportion", enabled_queue_order=t, enabled_reclaimable=t),
                                       PluginOption("nodeorder", enabled_node_order=t)]])
    yield "drf_first", PluginConf([[PluginOption("drf", enabled_job_order=t, enabled_preemptable=t), PluginOption("proportion", enabled_queue_order=t, enabled_reclaimable=t)],
                                   [PluginOption("gang", enabled_job_pipelined=t, enabled_preemptable=t, enabled_reclaimable=t), PluginOption("predicates", enabled_predicate=t),
                                    PluginOption("nodeorder", enabled_node_order=t, arguments={"mostrequested.weight": "2", "leastrequested.weight": "0"})]])
    yield "no_filters", PluginConf([[PluginOption("predicates", enabled_predicate=t), PluginOption("nodeorder", enabled_node_order=t)]])


def compare(what, o, ev, order, g, gev, gorder, states=None):
    np.testing.assert_array_equal(ev, gev, err_msg=f"{what}: evicted set")
    np.testing.assert_array_equal(order, gorder, err_msg=f"{what}: eviction order")
    for f in ("kind", "node", "step"):
        bad = np.nonzero(o.decisions[f] != g.decisions[f])[0]
        assert len(bad) == 0, f"{what}: decision field {f} differs at task {int(bad[0])}: oracle {o.decisions[bad[0]]} got {g.decisions[bad[0]]}"
    if states is not None:
        ns, os_ = states
        util.assert_same_state(o, ns, os_, what)


@pytest.mark.parametrize("case", list(golden_sessions()), ids=lambda c: c[0])
def test_reference_action_tests_on_the_emulation(case):
    name, action, s, tiers, expect = case
    o, ev, order = kbo.cycle(s, tiers, actions=(action,), running=s.meta["running"])
    assert int(ev.sum()) == expect                                   # the reference's expectation: number of cache.Evict calls
    g, gev, gorder = util.emu_evict(s, tiers, action, s.meta["running"])
    assert int(g.result.evictions) == expect
    compare(name, o, ev, order, g, gev, gorder, util.emu_states(g))


@pytest.mark.parametrize("seed", range(40))
def test_random_clusters_on_the_emulation(seed):
    s = random_cluster(seed)
    for tname, tiers in tier_variants():
        for action in ("reclaim", "preempt"):
            o, ev, order = kbo.cycle(s, tiers, actions=(action,), running=s.meta["running"])
            g, gev, gorder = util.emu_evict(s, tiers, action, s.meta["running"])
            compare(f"seed {seed} {tname} {action}", o, ev, order, g, gev, gorder, util.emu_states(g))
            assert int(g.result.evictions) == int(ev.sum())


@pytest.mark.parametrize("seed", range(3))
def test_larger_clusters_on_the_emulation(seed):
    s = random_cluster(100 + seed, big=True)
    for tname, tiers in (("default", PluginConf.default()),):
        for action in ("reclaim", "preempt"):
            o, ev, order = kbo.cycle(s, tiers, actions=(action,), running=s.meta["running"])
            g, gev, gorder = util.emu_evict(s, tiers, action, s.meta["running"])
            compare(f"big seed {seed} {tname} {action}", o, ev, order, g, gev, gorder, util.emu_states(g))


FULL_LIST = ("reclaim", "allocate", "backfill", "preempt")            # config/kube-batch-conf.yaml:1
ACTION_LISTS = (FULL_LIST, ("allocate", "backfill", "preempt"), ("reclaim", "allocate"), ("reclaim", "backfill", "preempt"), ("reclaim", "preempt"),
                ("allocate", "preempt", "preempt"))


@pytest.mark.parametrize("seed", range(30))
def test_action_lists_on_one_session_on_the_emulation(seed):
    """scheduler.go:88-101: the configured actions one after the other on ONE session — every action's own queues are filled from
    what the previous ones left (job / queue order keys, tasks still Pending, Running tasks not yet evicted)."""
    s = random_cluster(seed + 500)
    for tname, tiers in tier_variants():
        for acts in (ACTION_LISTS if seed % 3 == 0 else (FULL_LIST,)):
            for mode in ((1, 5) if seed % 2 else (1,)):
                o, ev, order = kbo.cycle(s, tiers, actions=acts, running=s.meta["running"])
                g, gev, gorder = util.emu_cycle(s, tiers, acts, s.meta["running"], mode=mode)
                what = f"seed {seed} {tname} {acts} mode {mode}"
                compare(what, o, ev, order, g, gev, gorder, util.emu_states(g))
                util.assert_same_decisions(o.decisions, g.decisions, what)


@pytest.mark.parametrize("seed", range(3))
def test_synthetic_cluster_with_a_thousand_evictions_on_the_emulation(seed):
    """BASELINE-shaped synthetic session (synth.py) + its Running filler pods one by one, 40 % of the filler PodGroups
    preemptable: the shipped action list evicts ~1000 pods, pipelines ~350 and allocates ~2000 tasks — all of it must match."""
    from kube_batch_b200 import synth
    s = synth.random_session(seed, tasks=3000, jobs=300, nodes=600, queues=3, oversub=2.0)
    run = synth.running_of(s, 0.4)
    o, ev, order = kbo.cycle(s, PluginConf.default(), actions=FULL_LIST, running=run)
    assert int(ev.sum()) > 500
    for mode in (1, 5):
        g, gev, gorder = util.emu_cycle(s, PluginConf.default(), FULL_LIST, run, mode=mode)
        compare(f"synthetic {seed} mode {mode}", o, ev, order, g, gev, gorder, util.emu_states(g))
        util.assert_same_decisions(o.decisions, g.decisions, f"synthetic {seed} mode {mode}")


def test_an_action_after_preempt_is_refused():
    """A discarded Statement leaves TaskInfo.NodeName behind (statement.go:153-188); a later ssn.Allocate of that task on another
    node fails in AddTask after the status change (node_info.go:173-176) — the oracle reproduces it, the engine refuses such
    action orders (the shipped order runs preempt last)."""
    s = random_cluster(518)
    with pytest.raises(RuntimeError, match="after preempt"):
        util.emu_cycle(s, PluginConf.default(), ("preempt", "allocate", "backfill"), s.meta["running"])


@pytest.mark.parametrize("seed", range(2))
def test_action_list_on_a_larger_cluster_on_the_emulation(seed):
    s = random_cluster(300 + seed, big=True)
    o, ev, order = kbo.cycle(s, PluginConf.default(), actions=FULL_LIST, running=s.meta["running"])
    g, gev, gorder = util.emu_cycle(s, PluginConf.default(), FULL_LIST, s.meta["running"], mode=5)
    compare(f"big {seed}", o, ev, order, g, gev, gorder, util.emu_states(g))
    util.assert_same_decisions(o.decisions, g.decisions, f"big {seed}")


# ------------------------------------------------------------------------------------------------ GPU, through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(2))
def test_synthetic_cluster_with_a_thousand_evictions_on_the_gpu(seed):
    from kube_batch_b200 import engine, synth
    s = synth.random_session(seed, tasks=3000, jobs=300, nodes=600, queues=3, oversub=2.0)
    run = synth.running_of(s, 0.4)
    o, ev, order = kbo.cycle(s, PluginConf.default(), actions=FULL_LIST, running=run)
    eng = engine.Engine(0)
    eng.load(s, PluginConf.default()).load_running(run)
    for rep in range(2):                                                    # repeatable from the loaded state
        res, gev, gorder, bounds = eng.cycle(FULL_LIST)
        compare(f"gpu synthetic {seed}", o, ev, order, res, gev, gorder, (eng.node_state(), eng.order_state()))
        util.assert_same_decisions(o.decisions, res.decisions, f"gpu synthetic {seed}")
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_action_lists_on_one_session_on_the_gpu(seed):
    from kube_batch_b200 import engine
    s = random_cluster(seed + 500, big=(seed % 4 == 3))
    eng = engine.Engine(0)
    for tname, tiers in tier_variants():
        eng.load(s, tiers).load_running(s.meta["running"])
        for acts in (ACTION_LISTS if seed % 2 == 0 else (FULL_LIST,)):
            o, ev, order = kbo.cycle(s, tiers, actions=acts, running=s.meta["running"])
            res, gev, gorder, bounds = eng.cycle(acts)
            what = f"gpu seed {seed} {tname} {acts}"
            compare(what, o, ev, order, res, gev, gorder, (eng.node_state(), eng.order_state()))
            util.assert_same_decisions(o.decisions, res.decisions, what)
            assert int(bounds[-1][1]) == int(ev.sum()) and (np.diff(bounds[:, 0].astype(np.int64)) >= 0).all()
        # kb_allocate from the loaded state is unaffected by the cycles before it (job lists / order slots restored)
        o1 = kbo.allocate(s, tiers)
        r1 = eng.allocate()
        util.assert_same_decisions(o1.decisions, r1.decisions, f"gpu seed {seed} {tname} allocate after cycles")
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(golden_sessions()), ids=lambda c: c[0])
def test_reference_action_tests_on_the_gpu(case):
    from kube_batch_b200 import engine
    name, action, s, tiers, expect = case
    o, ev, order = kbo.cycle(s, tiers, actions=(action,), running=s.meta["running"])
    eng = engine.Engine(0)
    eng.load(s, tiers).load_running(s.meta["running"])
    res, gev, gorder = eng.reclaim() if action == "reclaim" else eng.preempt()
    assert int(gev.sum()) == expect == int(res.stats.evictions)
    compare(name, o, ev, order, res, gev, gorder, (eng.node_state(), eng.order_state()))
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(10))
def test_random_clusters_on_the_gpu(seed):
    from kube_batch_b200 import engine
    s = random_cluster(seed, big=(seed % 3 == 2))
    eng = engine.Engine(0)
    for tname, tiers in tier_variants():
        eng.load(s, tiers).load_running(s.meta["running"])
        for action in ("reclaim", "preempt", "reclaim"):                     # repeatable: every action starts from the loaded state
            o, ev, order = kbo.cycle(s, tiers, actions=(action,), running=s.meta["running"])
            res, gev, gorder = eng.reclaim() if action == "reclaim" else eng.preempt()
            compare(f"gpu seed {seed} {tname} {action}", o, ev, order, res, gev, gorder, (eng.node_state(), eng.order_state()))
    eng.close()
