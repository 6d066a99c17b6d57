"""CPU-only parity: the engine's host/device-shared logic (kb_core.h / kb_ctl.h / kb_build.h), re-enacted
step by step by tests/emu (scan -> top-K -> certified replay -> control -> gang commit), must reproduce
the oracle bit-exactly.  The CUDA thread mechanics themselves are covered by the `-m gpu` tests."""
import numpy as np
import pytest

from kube_batch_b200 import builder as B
from kube_batch_b200 import synth
from kube_batch_b200.snapshot import PluginConf, PluginOption, Snapshot
from oracle import kbo
import util

CONFS = {
    "default": PluginConf.default(),
    "c1": synth.conf_c1(),
    "c2": synth.conf_c2(),
    "nogang": PluginConf.from_names([["priority"], ["drf", "predicates", "proportion", "nodeorder"]]),
    "none": PluginConf([]),
    "drf_first": PluginConf.from_names([["drf", "gang", "priority"], ["predicates", "nodeorder", "proportion"]]),
    "weights": PluginConf.from_names(
        [["priority", "gang"], ["drf", "predicates", "proportion", "nodeorder"]],
        {"nodeorder": {"leastrequested.weight": "0", "mostrequested.weight": "3", "balancedresource.weight": "2"},
         "predicates": {"predicate.MemoryPressureEnable": "true", "predicate.DiskPressureEnable": "true"}}),
    "negweight": PluginConf.from_names([["gang"], ["predicates", "nodeorder"]],
                                       {"nodeorder": {"leastrequested.weight": "-2", "balancedresource.weight": "1"}}),
    "allocate_test": PluginConf([[PluginOption("drf", enabled_preemptable=True, enabled_job_order=True),
                                  PluginOption("proportion", enabled_queue_order=True, enabled_reclaimable=True)]]),
}


def check(snap, conf, what):
    o = kbo.allocate(snap, conf)
    e = util.emu_allocate(snap, conf)
    util.assert_same_decisions(o.decisions, e.decisions, what)
    ns, os_ = util.emu_states(e)
    util.assert_same_state(o, ns, os_, what)
    assert e.result.tasks_processed == o.result.tasks_processed
    assert e.result.tasks_allocated == o.result.tasks_allocated
    assert e.result.tasks_pipelined == o.result.tasks_pipelined
    assert e.result.visits == o.result.visits
    assert e.result.jobs_ready == o.result.jobs_ready
    assert e.result.pairs_logical == o.result.pairs_logical
    return o, e


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_baseline_configs(name):
    s, conf = synth.make(name)
    check(s, conf, name)


@pytest.mark.parametrize("seed", range(24))
def test_random_sessions_all_confs(seed):
    rng = np.random.default_rng(seed)
    tasks = int(rng.integers(5, 300))
    jobs = int(rng.integers(1, min(tasks, 40) + 1))
    s = synth.random_session(seed, tasks=tasks, jobs=jobs, nodes=int(rng.integers(1, 200)), queues=int(rng.integers(1, 5)),
                             min_member_frac=float(rng.choice([0.0, 0.5, 1.0])), hetero=float(rng.choice([0, 0.3, 1.0])),
                             prio_levels=int(rng.integers(1, 4)), oversub=float(rng.choice([0.7, 1.3, 3.0])))
    for cname, conf in CONFS.items():
        check(s, conf, f"seed{seed}/{cname}")


@pytest.mark.parametrize("R,W", [(4, 3), (6, 4), (8, 4), (5, 2)])
def test_wide_records_more_dims_and_mask_words(R, W):
    # wider node records: extra scalar resources (R) and more label / taint / port mask words (W)
    for seed in range(3):
        s = synth.random_session(seed + 50, tasks=150, jobs=15, nodes=300, queues=2, hetero=0.3, R=R, W=W)
        for cname in ("default", "c2"):
            check(s, CONFS[cname], f"R{R}W{W}/seed{seed}/{cname}")


def test_long_run_forces_rescans():
    # one job, 400 identical tasks, spreading score: > DMAX distinct nodes get dirtied inside one run
    s = synth.random_session(7, tasks=400, jobs=1, nodes=300, hetero=0.0, oversub=0.5)
    o, e = check(s, synth.conf_c2(), "long-run")
    assert e.result.kernel_launches > 5 and o.result.tasks_allocated > 100


def test_empty_and_degenerate_sessions():
    for (T, J, N) in [(0, 0, 0), (0, 0, 5), (3, 1, 0)]:
        s = Snapshot(3, 1, N, T, J, 1)
        s.job_task_off[:] = [0] + [T] * J
        s.task_uid_rank[:] = np.arange(T)
        s.task_resreq[0, :] = 1000
        s.task_initreq[0, :] = 1000
        s.queue_weight[:] = 1
        if N:
            s.node_idle[0, :] = 4000
            s.node_allocatable[0, :] = 4000
            s.node_max_pods[:] = 10
        check(s, PluginConf.default(), f"degenerate T{T} J{J} N{N}")


def test_pipeline_onto_releasing():
    # node full but a pod is terminating: the task must be Pipelined (allocate.go:175-181), never dispatched
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "old", "q", min_member=1))
    b.add_pod_group(B.PodGroup("ns", "new", "q", min_member=1))
    b.add_node(B.build_node("n1", {"cpu": 4, "memory": 8e9}, pods=10))
    b.add_pod(B.Pod("ns", "dying", "n1", "Running", {"cpu": 4, "memory": 8e9}, group="old", deleting=True))
    b.add_pod(B.Pod("ns", "p", "", "Pending", {"cpu": 2, "memory": 1e9}, group="new"))
    s = b.flatten()
    o, e = check(s, PluginConf.default(), "pipeline")
    assert int(o.decisions["kind"][0]) == 2 and int(o.decisions["dispatched"][0]) == 0
    assert o.node_releasing[0, 0] == 2000.0


def test_host_ports_taints_selectors_affinity():
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "g", "q", min_member=1))
    for i, (zone, taint) in enumerate([("a", None), ("b", ("dedicated", "batch", "NoSchedule")), ("b", None), ("c", None)]):
        b.add_node(B.Node(f"n{i}", {"cpu": 8, "memory": 32e9, "pods": 10}, labels={"zone": zone, "rank": str(i)},
                          taints=[taint] if taint else []))
    req = {"cpu": 1, "memory": 1e9}
    b.add_pod(B.Pod("ns", "sel-b", "", "Pending", req, group="g", node_selector={"zone": "b"}, creation=1))
    b.add_pod(B.Pod("ns", "sel-b-tol", "", "Pending", req, group="g", node_selector={"zone": "b"},
                    tolerations=[("dedicated", "Equal", "batch", "NoSchedule")], creation=2))
    b.add_pod(B.Pod("ns", "port-1", "", "Pending", req, group="g", host_ports=[("", "TCP", 8080)], creation=3))
    b.add_pod(B.Pod("ns", "port-2", "", "Pending", req, group="g", host_ports=[("10.0.0.1", "TCP", 8080)], creation=4))
    b.add_pod(B.Pod("ns", "aff", "", "Pending", req, group="g", creation=5,
                    affinity_terms=[[("zone", "In", ["c"])], [("rank", "Gt", ["2"]), ("zone", "NotIn", ["a"])]]))
    b.add_pod(B.Pod("ns", "nowhere", "", "Pending", req, group="g", node_selector={"zone": "z"}, creation=6))
    s = b.flatten()
    conf = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]])
    o, e = check(s, conf, "predicates")
    got = {s.meta["tasks"][t]: (s.meta["nodes"][int(o.decisions["node"][t])] if o.decisions["node"][t] >= 0 else None)
           for t in range(s.T)}
    assert got["ns/sel-b"] == "n2"            # n1 is tainted
    assert got["ns/sel-b-tol"] == "n1"        # least-requested prefers the still-empty tolerated node
    assert got["ns/port-1"] != got["ns/port-2"] and got["ns/port-1"] is not None   # wildcard IP conflicts with 10.0.0.1
    assert got["ns/aff"] == "n3"
    assert got["ns/nowhere"] is None


def test_unknown_plugin_is_refused():
    s, _ = synth.make("c1")
    bad = PluginConf.from_names([["gang", "my-custom-plugin"]])
    with pytest.raises(RuntimeError):
        util.emu_allocate(s, bad)
    with pytest.raises(RuntimeError):
        kbo.allocate(s, bad)
