"""Pins the CPU oracle (oracle/kb_oracle.cpp) against the reference's OWN golden vectors.

Every expected value below is transcribed from a test in /root/reference (file:line cited per case);
the reference is Go and cannot run here, so these known answers are the pin.
"""
import numpy as np
import pytest

from kube_batch_b200 import builder as B
from kube_batch_b200.snapshot import PluginConf, PluginOption
from kube_batch_b200 import abi
from oracle import kbo

S1, S2 = 2, 3          # dims: 0 cpu, 1 memory, 2 "scalar.test/scalar1", 3 "hugepages-test"
P12 = (1 << S1) | (1 << S2)
R4 = 4


# ---- pkg/scheduler/api/resource_info_test.go:246-304 TestLessEqual ----
@pytest.mark.parametrize("l,lp,r,rp,exp", [
    ([0, 0, 0, 0], 0, [4000, 2000, 1000, 2000], P12, True),
    ([4000, 4000, 1000, 2000], P12, [2000, 2000, 4000, 5000], P12, False),
    ([4, 4000, 1, 0], 1 << S1, [0, 0, 0, 0], 0, True),            # the epsilon case (:274-282)
    ([4000, 4000, 1000, 2000], P12, [8000, 8000, 4000, 5000], P12, True),
])
def test_less_equal_golden(l, lp, r, rp, exp):
    assert kbo.res_less_equal(l, lp, r, rp, R=R4) is exp


# ---- resource_info_test.go:352-419 TestLess ----
@pytest.mark.parametrize("l,lp,r,rp,exp", [
    ([0, 0, 0, 0], 0, [0, 0, 0, 0], 0, False),
    ([0, 0, 0, 0], 0, [4000, 2000, 1000, 2000], P12, True),
    ([4000, 4000, 1000, 2000], P12, [8000, 8000, 4000, 5000], P12, True),
    ([4000, 4000, 5000, 2000], P12, [8000, 8000, 4000, 5000], P12, False),
    ([9000, 4000, 1000, 2000], P12, [8000, 8000, 4000, 5000], P12, False),
])
def test_less_golden(l, lp, r, rp, exp):
    assert kbo.res_less(l, lp, r, rp, R=R4) is exp


# ---- resource_info_test.go:306-350 TestSubResource ----
def test_sub_golden():
    v, p = kbo.res_sub([4000, 2000, 1, 2], P12, [0, 0, 0, 0], 0, R=R4)
    assert list(v) == [4000, 2000, 1, 2] and p == P12
    v, p = kbo.res_sub([4000, 4000, 1000, 2000], P12, [3000, 2000, 500, 1000], P12, R=R4)
    assert list(v) == [1000, 2000, 500, 1000] and p == P12
    with pytest.raises(ArithmeticError):        # resource_info.go:158 panics
        kbo.res_sub([1000, 1000, 0, 0], 0, [3000, 2000, 0, 0], 0, R=R4)


# ---- resource_info_test.go:183-244 TestAddResource ----
def test_add_golden():
    v, p = kbo.res_add([0, 0, 0, 0], 0, [4000, 2000, 1, 2], P12, R=R4)
    assert list(v) == [4000, 2000, 1, 2] and p == P12
    v, p = kbo.res_add([4000, 4000, 1, 2], P12, [4000, 2000, 4, 5], P12, R=R4)
    assert list(v) == [8000, 6000, 5, 7] and p == P12
    v, p = kbo.res_add([4000, 4000, 1, 0], 1 << S1, [4000, 2000, 4, 5], P12, R=R4)
    assert list(v) == [8000, 6000, 5, 5] and p == P12


# ---- resource_info_test.go:98-142 TestSetMaxResource ----
def test_set_max_golden():
    v, p = kbo.res_set_max([0, 0, 0, 0], 0, [4000, 2000, 1, 2], P12, R=R4)
    assert list(v) == [4000, 2000, 1, 2] and p == P12
    v, p = kbo.res_set_max([4000, 4000, 1, 2], P12, [4000, 2000, 4, 5], P12, R=R4)
    assert list(v) == [4000, 4000, 4, 5] and p == P12


def test_is_empty_thresholds():
    # resource_info.go:93-105 with mins :68-70
    assert kbo.res_is_empty([9, 10 * 1024 * 1024 - 1, 9, 0], 1 << S1, R=R4)
    assert not kbo.res_is_empty([10, 0, 0, 0], 0, R=R4)
    assert not kbo.res_is_empty([0, 10 * 1024 * 1024, 0, 0], 0, R=R4)
    assert not kbo.res_is_empty([0, 0, 10, 0], 1 << S1, R=R4)


# ---- pkg/scheduler/api/node_info_test.go:35-68: 8000m/10G node, running pods 1000m/1G + 2000m/2G ----
def test_node_info_add_pod_golden():
    b = B.SessionBuilder()
    b.add_node(B.build_node("n1", {"cpu": 8, "memory": 10e9}, pods=110))
    b.add_queue(B.Queue("q"))
    b.add_pod_group(B.PodGroup("c1", "pg", "q"))
    b.add_pod(B.Pod("c1", "p1", "n1", "Running", {"cpu": 1, "memory": 1e9}, group="pg"))
    b.add_pod(B.Pod("c1", "p2", "n1", "Running", {"cpu": 2, "memory": 2e9}, group="pg"))
    s = b.flatten()
    assert s.node_idle[:2, 0].tolist() == [5000.0, 7e9]
    assert s.node_used[:2, 0].tolist() == [3000.0, 3e9]
    assert s.node_releasing[:2, 0].tolist() == [0.0, 0.0]
    assert s.node_pods[0] == 2 and s.job_ready0[0] == 2


# ---- pkg/scheduler/api/pod_info_test.go:53-86: init-container max rule -> 3000m / 5G ----
def test_pod_init_container_golden():
    b = B.SessionBuilder()
    b.add_node(B.build_node("n1", {"cpu": 8, "memory": 10e9}, pods=110))
    b.add_queue(B.Queue("q"))
    b.add_pod_group(B.PodGroup("c1", "pg", "q"))
    # two containers 1000m/1G + 2000m/1G are summed by the caller into one request list
    b.add_pod(B.Pod("c1", "p1", "", "Pending", {"cpu": 3, "memory": 2e9},
                    init_requests=[{"cpu": 2, "memory": 5e9}, {"cpu": 2, "memory": 1e9}], group="pg"))
    s = b.flatten()
    assert s.task_initreq[:2, 0].tolist() == [3000.0, 5e9]
    assert s.task_resreq[:2, 0].tolist() == [3000.0, 2e9]


def _tiers_allocate_test():
    # allocate_test.go:180-195
    return PluginConf([[PluginOption("drf", enabled_preemptable=True, enabled_job_order=True),
                        PluginOption("proportion", enabled_queue_order=True, enabled_reclaimable=True)]])


# ---- pkg/scheduler/actions/allocate/allocate_test.go:51-85 "one Job with two Pods on one node" ----
def test_allocate_case1_golden():
    b = B.SessionBuilder()
    b.add_pod_group(B.PodGroup("c1", "pg1", "c1"))
    b.add_pod(B.build_pod("c1", "p1", "", "Pending", B.build_resource_list("1", "1G"), "pg1"))
    b.add_pod(B.build_pod("c1", "p2", "", "Pending", B.build_resource_list("1", "1G"), "pg1"))
    b.add_node(B.build_node("n1", B.build_resource_list("2", "4Gi")))
    b.add_queue(B.Queue("c1", 1))
    s = b.flatten()
    o = kbo.allocate(s, _tiers_allocate_test())
    binds = {s.meta["tasks"][t]: s.meta["nodes"][n] for t, n in o.bind_map().items()}
    assert binds == {"c1/p1": "n1", "c1/p2": "n1"}


# ---- allocate_test.go:86-144 "two Jobs on one node" ----
def test_allocate_case2_golden():
    b = B.SessionBuilder()
    b.add_pod_group(B.PodGroup("c1", "pg1", "c1"))
    b.add_pod_group(B.PodGroup("c2", "pg2", "c2"))
    for ns, pg in (("c1", "pg1"), ("c2", "pg2")):
        b.add_pod(B.build_pod(ns, "p1", "", "Pending", B.build_resource_list("1", "1G"), pg))
        b.add_pod(B.build_pod(ns, "p2", "", "Pending", B.build_resource_list("1", "1G"), pg))
    b.add_node(B.build_node("n1", B.build_resource_list("2", "4G")))
    b.add_queue(B.Queue("c1", 1))
    b.add_queue(B.Queue("c2", 1))
    s = b.flatten()
    o = kbo.allocate(s, _tiers_allocate_test())
    binds = {s.meta["tasks"][t]: s.meta["nodes"][n] for t, n in o.bind_map().items()}
    assert binds == {"c2/p1": "n1", "c1/p1": "n1"}
    # proportion: deserved = (1000 m, 2e9) each (hand trace in SURVEY.md §8c)
    assert o.queue_deserved[:2, 0].tolist() == [1000.0, 2e9]
    assert o.queue_deserved[:2, 1].tolist() == [1000.0, 2e9]


# ---- fixture gotcha: util.BuildNode sets no `pods` => predicates rejects every node ----
def test_predicates_rejects_nodes_without_pod_capacity():
    b = B.SessionBuilder()
    b.add_pod_group(B.PodGroup("c1", "pg1", "c1"))
    b.add_pod(B.build_pod("c1", "p1", "", "Pending", B.build_resource_list("1", "1G"), "pg1"))
    b.add_node(B.build_node("n1", B.build_resource_list("2", "4Gi")))
    b.add_queue(B.Queue("c1", 1))
    s = b.flatten()
    conf = PluginConf([[PluginOption("predicates", enabled_predicate=True)]])
    assert kbo.allocate(s, conf).bind_map() == {}


# ---- hand-computed priority vectors: vendored k8s arithmetic has no surviving tests ----
def test_priority_scores_hand_computed():
    GiB = 1 << 30
    # alloc 4000m/8GiB, node nz 1000m/2GiB, pod 500m/1GiB
    assert kbo.lib().kbo_least_requested(1500, 4000, 3 * GiB, 8 * GiB) == 6
    assert kbo.lib().kbo_balanced(1500, 4000, 3 * GiB, 8 * GiB) == 10
    assert kbo.lib().kbo_most_requested(1500, 4000, 3 * GiB, 8 * GiB) == 3
    # requested > capacity and capacity == 0 (least_requested.go:49-58, balanced_resource_allocation.go:42-79)
    assert kbo.lib().kbo_least_requested(5000, 4000, 9 * GiB, 8 * GiB) == 0
    assert kbo.lib().kbo_balanced(4000, 4000, 1 * GiB, 8 * GiB) == 0
    assert kbo.lib().kbo_least_requested(0, 0, 0, 0) == 0
    assert kbo.lib().kbo_balanced(0, 0, 0, 8 * GiB) == 0      # fraction(…, 0) = 1 -> >= 1 -> 0
    # (1 - |0.25 - 0.75|) * 10 = 5
    assert kbo.lib().kbo_balanced(1000, 4000, 6 * GiB, 8 * GiB) == 5


def test_share_golden():
    # api/helpers/helpers.go:47-60
    assert kbo.lib().kbo_share(0.0, 0.0) == 0.0
    assert kbo.lib().kbo_share(5.0, 0.0) == 1.0
    assert kbo.lib().kbo_share(1.0, 4.0) == 0.25


def test_heap_is_container_heap():
    # go1.13 container/heap pops a strict total order sorted, whatever the push order
    rng = np.random.default_rng(0)
    k = rng.permutation(257)
    assert kbo.heap_sort(k).tolist() == sorted(k.tolist())


# ---- gang semantics (plugins/gang/gang.go:122-125, framework/session.go:277-285) ----
def test_gang_dispatch_only_when_min_member_reached():
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "fits", "q", min_member=2))
    b.add_pod_group(B.PodGroup("ns", "toobig", "q", min_member=3))
    for i in range(2):
        b.add_pod(B.build_pod("ns", f"a{i}", "", "Pending", B.build_resource_list("1", "1G"), "fits"))
    for i in range(3):
        b.add_pod(B.build_pod("ns", f"b{i}", "", "Pending", B.build_resource_list("1", "1G"), "toobig"))
    b.add_node(B.build_node("n1", B.build_resource_list("4", "16G"), pods=110))
    s = b.flatten()
    conf = PluginConf([[PluginOption("gang", enabled_job_order=True, enabled_job_ready=True)]])
    o = kbo.allocate(s, conf)
    names = s.meta["tasks"]
    d = o.decisions
    got = {names[t]: (int(d["kind"][t]), int(d["dispatched"][t])) for t in range(s.T)}
    # "fits" is served first (ascending JobID) and dispatched; "toobig" gets 2 of 3 tasks Allocated that
    # hold node.Idle for the rest of the cycle but are never dispatched (no rollback in allocate)
    assert got["ns/a0"] == (1, 1) and got["ns/a1"] == (1, 1)
    assert sorted(v for k, v in got.items() if k.startswith("ns/b")) == [(0, 0), (1, 0), (1, 0)]
    assert o.node_idle[0, 0] == 0.0


# ---------------- NodeAffinityPriority: vendor/.../priorities/node_affinity.go:34-77 + reduce.go:28-63 ----------------
# The reference carries no test for it ("parity unpinned by the reference"); the vectors below are hand-computed from the
# vendored code.  The ENGINE refuses such tasks in this build (KB_E_UNSUPPORTED_FEATURE) — this pins the oracle for the next round.
def _pref_session(pods):
    from kube_batch_b200 import builder as B
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "g", "q", min_member=1))
    for name, labels, taints in (("n0", {"zone": "a"}, []), ("n1", {"zone": "b", "disk": "ssd"}, [("dedicated", "x", "NoSchedule")]),
                                 ("n2", {"zone": "b"}, []), ("n3", {"zone": "c"}, [])):
        b.add_node(B.Node(name, {"cpu": 8, "memory": 32e9, "pods": 10}, labels=labels, taints=taints))
    for p in pods:
        b.add_pod(p)
    return b.flatten()


def test_node_affinity_priority_counts_and_normalisation():
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf
    req = {"cpu": 1, "memory": 1e9}
    pref = [(80, [("zone", "In", ["b"])]), (20, [("disk", "In", ["ssd"])]), (0, [("zone", "In", ["c"])])]
    tol = [("dedicated", "Equal", "x", "NoSchedule")]
    s = _pref_session([B.Pod("ns", f"p{i}", "", "Pending", req, group="g", creation=i, preferred_terms=pref, tolerations=tol) for i in range(3)])
    conf = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]])
    # counts: n0 0, n1 100, n2 80, n3 0 (the weight-0 term is skipped) -> NormalizeReduce(10): 0, 10, 8, 0
    fit, score = kbo.predicate_score(s, conf, 0)
    # base = least (8+9)/2 = 8 + balanced int((1 - |0.125 - 0.03125|) * 10) = 9 -> 17 on every (empty, identical) node
    assert fit.tolist() == [1, 1, 1, 1] and score.tolist() == [17.0, 27.0, 25.0, 17.0]
    o = kbo.allocate(s, conf)
    names = [s.meta["nodes"][n] for n in o.decisions["node"]]
    # p0 -> n1 (27).  p1: n1 now scores least 8 + balanced 8 + 10 = 26 > n2 25.  p2: n1 7 + 7 + 10 = 24 < n2 25.
    assert names == ["n1", "n1", "n2"]


def test_node_affinity_priority_normalises_over_feasible_nodes_only():
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf
    req = {"cpu": 1, "memory": 1e9}
    pref = [(80, [("zone", "In", ["b"])]), (20, [("disk", "In", ["ssd"])])]
    # no toleration: n1 (the only 100-count node) is infeasible, so maxCount = 80 and n2 gets the full 10
    s = _pref_session([B.Pod("ns", "p", "", "Pending", req, group="g", preferred_terms=pref)])
    conf = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "3"}})
    fit, score = kbo.predicate_score(s, conf, 0)
    assert fit.tolist() == [1, 0, 1, 1] and score[[0, 2, 3]].tolist() == [17.0, 17.0 + 3 * 10, 17.0]
    o = kbo.allocate(s, conf)
    assert s.meta["nodes"][int(o.decisions["node"][0])] == "n2"
    # an empty preferred term matches every node: all counts equal -> every node gets 10 -> no effect on the order
    s = _pref_session([B.Pod("ns", "p", "", "Pending", req, group="g", preferred_terms=[(5, [])])])
    fit, score = kbo.predicate_score(s, PluginConf.from_names([["gang"], ["predicates", "nodeorder"]]), 0)
    assert score[[0, 2, 3]].tolist() == [27.0, 27.0, 27.0]
    # nodeaffinity.weight 0 switches the term off
    s = _pref_session([B.Pod("ns", "p", "", "Pending", req, group="g", preferred_terms=pref)])
    fit, score = kbo.predicate_score(s, PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "0"}}), 0)
    assert score[[0, 2, 3]].tolist() == [17.0, 17.0, 17.0]


def test_sub_milli_quantities_round_up_like_milli_value():
    """resource.Quantity.MilliValue() rounds up: cpu 0.0005 is 1 milli (api.NewResource, resource_info.go:73-90, and
    GetNonzeroRequests), not 0 as half-to-even rounding would give."""
    from kube_batch_b200 import builder as B
    v, _ = B.SessionBuilder._resource({"cpu": 0.0005, "memory": 1.0}, ["cpu", "memory"])
    assert v.tolist() == [1.0, 1.0]
    assert B.SessionBuilder._milli(0.5) == 500 and B.SessionBuilder._milli(2.5e-3) == 3 and B.SessionBuilder._milli(1.0) == 1000


def test_per_launch_kernels_evaluate_preferred_node_affinity_on_the_counter_path():
    """NodeAffinityPriority runs in cycle_kernel (two-pass scan); a session outside the pipeline's record geometry — here R = 2,
    W = 1 — runs on the per-visit kernels, which evaluate it with the inter-pod machinery (a pass over the feasible nodes before the
    visit for the max count, a fresh scan per task of such a class).  Round 1 / early round 2 refused such sessions."""
    import util
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf
    from oracle import kbo
    s = _pref_session([B.Pod("ns", "p", "", "Pending", {"cpu": 1}, group="g", preferred_terms=[(1, [("zone", "In", ["a"])])])])
    assert s.R == 2
    o = kbo.allocate(s, PluginConf.default())
    for mode in (0, 1, 5):
        e = util.emu_allocate(s, PluginConf.default(), mode=mode)
        util.assert_same_decisions(o.decisions, e.decisions, f"mode {mode}")


# ---------------- preempt / reclaim: ORACLE ONLY so far; pinned on the reference's own action tests ----------------
def _evict_session(pods, node_res, queues, groups):
    from kube_batch_b200 import builder as B
    b = B.SessionBuilder()
    for g, q in groups:
        b.add_pod_group(B.PodGroup("c1", g, q))
    for name, node, phase, group in pods:
        b.add_pod(B.build_pod("c1", name, node, phase, B.build_resource_list("1", "1G"), group))
    b.add_node(B.build_node("n1", B.build_resource_list(*node_res)))
    for q in queues:
        b.add_queue(B.Queue(q, 1))
    return b.flatten()


def test_preempt_golden():
    """preempt_test.go:51-144 — tiers conformance + gang with EnabledPreemptable; expected = number of cache.Evict calls."""
    from kube_batch_b200.snapshot import PluginConf, PluginOption
    tiers = PluginConf([[PluginOption("conformance", enabled_preemptable=True), PluginOption("gang", enabled_preemptable=True)]])
    # "one Job with two Pods on one node" (:51-85): expected 1
    s = _evict_session([("preemptee1", "n1", "Running", "pg1"), ("preemptee2", "n1", "Running", "pg1"),
                        ("preemptor1", "", "Pending", "pg1"), ("preemptor2", "", "Pending", "pg1")], ("3", "3Gi"), ["q1"], [("pg1", "q1")])
    o, ev, order = kbo.cycle(s, tiers, actions=("preempt",), running=s.meta["running"])
    assert o.result.evictions == 1 and int(ev.sum()) == 1
    assert (o.decisions["kind"] == 2).sum() == 1          # one preemptor pipelined onto the freed resources
    # "two Jobs on one node" (:86-144): expected 2
    s = _evict_session([("preemptee1", "n1", "Running", "pg1"), ("preemptee2", "n1", "Running", "pg1"),
                        ("preemptor1", "", "Pending", "pg2"), ("preemptor2", "", "Pending", "pg2")], ("2", "2G"), ["q1"],
                       [("pg1", "q1"), ("pg2", "q1")])
    o, ev, order = kbo.cycle(s, tiers, actions=("preempt",), running=s.meta["running"])
    assert o.result.evictions == 2 and ev.all() and sorted(order.tolist()) == [0, 1]
    assert (o.decisions["kind"] == 2).all() and not o.decisions["dispatched"].any()


def test_reclaim_golden():
    """reclaim_test.go:51-105 — two queues, q1 holds the whole node, a pending task of q2 reclaims one pod: expected 1."""
    from kube_batch_b200.snapshot import PluginConf, PluginOption
    tiers = PluginConf([[PluginOption("conformance", enabled_reclaimable=True), PluginOption("gang", enabled_reclaimable=True)]])
    s = _evict_session([("preemptee1", "n1", "Running", "pg1"), ("preemptee2", "n1", "Running", "pg1"), ("preemptee3", "n1", "Running", "pg1"),
                        ("preemptor1", "", "Pending", "pg2")], ("3", "3Gi"), ["q1", "q2"], [("pg1", "q1"), ("pg2", "q2")])
    o, ev, order = kbo.cycle(s, tiers, actions=("reclaim",), running=s.meta["running"])
    assert o.result.evictions == 1 and ev.tolist() == [True, False, False]      # node.Tasks in UID order, first victim suffices
    assert int(o.decisions["kind"][0]) == 2 and int(o.decisions["node"][0]) == 0
    assert o.node_releasing[0, 0] == 0.0 and o.node_used[0, 0] == 4000.0          # evicted pod still counted in Used + the pipelined one


def test_gang_protects_min_available_and_priority_orders_victims():
    """Hand-computed: gang refuses victims whose job would drop below MinAvailable (gang.go:70-90); among the victims the
    lowest-priority task goes first (preempt.go:210-231); a discarded statement leaves no trace (statement.go:191-203)."""
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf, PluginOption
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "held", "q", min_member=2))      # 2 running, MinAvailable 2 -> not preemptable (2 <= 2-1 is false)
    b.add_pod_group(B.PodGroup("ns", "loose", "q", min_member=1))     # MinAvailable == 1 -> preemptable
    b.add_pod_group(B.PodGroup("ns", "want", "q", min_member=1))
    b.add_node(B.Node("n0", {"cpu": 4, "memory": 16e9, "pods": 10}))
    req = {"cpu": 1, "memory": 1e9}
    b.add_pod(B.Pod("ns", "h0", "n0", "Running", req, group="held"))
    b.add_pod(B.Pod("ns", "h1", "n0", "Running", req, group="held"))
    b.add_pod(B.Pod("ns", "l-hi", "n0", "Running", req, group="loose", priority=9))
    b.add_pod(B.Pod("ns", "l-lo", "n0", "Running", req, group="loose", priority=1))
    b.add_pod(B.Pod("ns", "w0", "", "Pending", req, group="want"))
    b.add_pod(B.Pod("ns", "w-big", "", "Pending", {"cpu": 3, "memory": 1e9}, group="want", creation=5))
    s = b.flatten()
    tiers = PluginConf([[PluginOption("priority", enabled_task_order=True), PluginOption("gang", enabled_preemptable=True, enabled_job_pipelined=True),
                         PluginOption("conformance", enabled_preemptable=True)]])
    o, ev, order = kbo.cycle(s, tiers, actions=("preempt",), running=s.meta["running"])
    names = s.meta["running"]["names"]
    evicted = {names[i] for i in range(len(names)) if ev[i]}
    # w0 (created first) preempts the lowest-priority loose pod; w-big needs 3 cpu but only l-hi (1 cpu) is left as a victim -> nothing
    assert evicted == {"ns/l-lo"}
    t = {s.meta["tasks"][i]: i for i in range(s.T)}
    assert int(o.decisions["kind"][t["ns/w0"]]) == 2 and int(o.decisions["kind"][t["ns/w-big"]]) == 0
    assert int(o.decisions["node"][t["ns/w-big"]]) == -1


@pytest.mark.parametrize("seed", range(12))
def test_full_action_list_keeps_the_books_balanced(seed):
    """"reclaim, allocate, backfill, preempt" (config/kube-batch-conf.yaml:1) on random clusters with Running, terminating and
    Pending pods in several queues: whatever the actions decide, the node / eviction bookkeeping must stay consistent."""
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf
    rng = np.random.default_rng(4000 + seed)
    b = B.SessionBuilder()
    nq = int(rng.integers(1, 4))
    for q in range(nq):
        b.add_queue(B.Queue(f"q{q}", int(rng.integers(1, 4))))
    nn = int(rng.integers(2, 9))
    for n in range(nn):
        b.add_node(B.Node(f"n{n}", {"cpu": 8, "memory": 32e9, "pods": 12}))
    cap = {f"n{n}": 8.0 for n in range(nn)}
    pods = []
    for g in range(int(rng.integers(2, 9))):
        b.add_pod_group(B.PodGroup("ns", f"g{g}", f"q{int(rng.integers(0, nq))}", min_member=int(rng.integers(0, 4)), priority=int(rng.integers(0, 3))))
        cpu = float(rng.choice([0.5, 1, 2]))
        for k in range(int(rng.integers(1, 7))):
            state = rng.choice(["Running", "Running", "Pending", "Pending", "Deleting"])
            node = ""
            if state != "Pending":
                free = [n for n, c in cap.items() if c >= cpu]
                if not free:
                    state = "Pending"
                else:
                    node = str(rng.choice(free))
                    cap[node] -= cpu
            p = B.Pod("ns", f"g{g}-p{k}", node, "Pending" if state == "Pending" else "Running", {"cpu": cpu, "memory": cpu * 1e9},
                      group=f"g{g}", priority=int(rng.integers(0, 3)), creation=len(pods), deleting=(state == "Deleting"))
            pods.append(p)
            b.add_pod(p)
    s = b.flatten()
    o, ev, order = kbo.cycle(s, PluginConf.default(), actions=("reclaim", "allocate", "backfill", "preempt"), running=s.meta["running"])
    k = int(ev.sum())
    assert o.result.evictions == k and sorted(order[ev].tolist()) == list(range(k))
    # per node: Allocatable = Idle + Used - (pipelined tasks); Releasing = terminating + evicted - pipelined >= 0
    rt = s.meta["running"]
    pip = np.zeros((s.R, s.N)); rel = np.zeros((s.R, s.N))
    d = o.decisions
    for t in np.nonzero(d["kind"] == abi.KB_KIND_PIPELINED)[0]:
        pip[:, d["node"][t]] += s.task_resreq[:, t]
    for i in np.nonzero(ev)[0]:
        rel[:, rt["node"][i]] += rt["resreq"][:, i]
    for p in pods:
        if p.deleting:
            n = s.meta["nodes"].index(p.node_name)
            rel[0, n] += p.requests["cpu"] * 1000
            rel[1, n] += p.requests["memory"]
    np.testing.assert_allclose(o.node_idle + o.node_used - pip, s.node_allocatable, rtol=0, atol=1e-3)
    np.testing.assert_allclose(o.node_releasing, rel - pip, rtol=0, atol=1e-3)
    assert (o.node_idle >= -1e-6).all() and (o.node_releasing >= -1e-6).all()
    placed = d["kind"] != abi.KB_KIND_NONE
    assert (d["node"][(d["kind"] == abi.KB_KIND_ALLOCATED) | (d["kind"] == abi.KB_KIND_PIPELINED)] >= 0).all()
    assert (d["node"][d["kind"] == abi.KB_KIND_NONE] == -1).all() and placed.sum() >= 0


def test_proportion_reclaims_only_what_exceeds_the_deserved_share():
    """Hand-computed (proportion.go:171-196): a reclaimee is a victim only while its queue's allocation, minus everything already
    counted, stays >= deserved IN EVERY DIMENSION.  Two queues of weight 1 on one node of 8 cpu / 8 GB: deserved = 4 cpu / 4 GB each
    (both request more).  q-rich runs 6 x (1 cpu, 1 GB): only the first two reclaimees (allocated 6 -> 5 -> 4) are victims."""
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf, PluginOption

    def session(cpu):
        b = B.SessionBuilder()
        b.add_queue(B.Queue("q-poor", 1))
        b.add_queue(B.Queue("q-rich", 1))
        b.add_pod_group(B.PodGroup("ns", "rich", "q-rich", min_member=1))
        b.add_pod_group(B.PodGroup("ns", "poor", "q-poor", min_member=1))
        b.add_node(B.Node("n0", {"cpu": 8, "memory": 8e9, "pods": 20}))
        for i in range(6):
            b.add_pod(B.Pod("ns", f"r{i}", "n0", "Running", {"cpu": 1, "memory": 1e9}, group="rich"))
        for i in range(4):
            b.add_pod(B.Pod("ns", f"p{i}", "", "Pending", {"cpu": cpu, "memory": cpu * 1e9}, group="poor", creation=i))
        return b.flatten()

    tiers = PluginConf([[PluginOption("gang", enabled_reclaimable=True), PluginOption("proportion", enabled_reclaimable=True, enabled_queue_order=True)]])
    s = session(3)
    o, ev, order = kbo.cycle(s, tiers, actions=("reclaim",), running=s.meta["running"])
    # water-filling: total 8 cpu; requests rich 6, poor 12 -> 4 / 4, nothing left
    np.testing.assert_array_equal(o.queue_deserved[0], [4000.0, 4000.0])
    # p0 needs 3 cpu: the victims on n0 are r0, r1 only (2 cpu) -> "not enough resource from victims": nothing evicted
    assert o.result.evictions == 0 and (o.decisions["kind"] == 0).all()
    # a 2-cpu task: the two victims suffice; both are evicted, the task is pipelined; reclaim pops ONE task per job visit and does
    # not push the job back, so p1..p3 are never tried
    s = session(2)
    o, ev, order = kbo.cycle(s, tiers, actions=("reclaim",), running=s.meta["running"])
    names = s.meta["running"]["names"]
    assert {names[i] for i in np.nonzero(ev)[0]} == {"ns/r0", "ns/r1"} and o.result.evictions == 2
    assert (o.decisions["kind"] == 2).sum() == 1 and int(o.decisions["kind"][0]) == 2
    np.testing.assert_array_equal(o.queue_allocated[0], [2000.0, 4000.0])      # q-poor got 2 cpu pipelined, q-rich gave 2 back


def test_drf_preempts_only_towards_a_fairer_split():
    """Hand-computed (drf.go:84-110): preemptee is a victim iff the preemptor job's share AFTER gaining its task is <= the preemptee
    job's share AFTER losing that task (cumulatively per job).  10-cpu node: job big runs 6 x 1 cpu (share 0.6), job small runs 1
    (0.1) and wants one more: ls = 0.2; big: 0.5 (victim), 0.4 (victim), ... 0.2 -> victim while rs >= 0.2 - 1e-6: four tasks."""
    from kube_batch_b200 import builder as B
    from kube_batch_b200.snapshot import PluginConf, PluginOption
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "big", "q", min_member=1))
    b.add_pod_group(B.PodGroup("ns", "small", "q", min_member=1))
    b.add_node(B.Node("n0", {"cpu": 10, "memory": 100e9, "pods": 20}))
    for i in range(6):
        b.add_pod(B.Pod("ns", f"b{i}", "n0", "Running", {"cpu": 1, "memory": 1e9}, group="big"))
    b.add_pod(B.Pod("ns", "s-run", "n0", "Running", {"cpu": 1, "memory": 1e9}, group="small"))
    b.add_pod(B.Pod("ns", "s-new", "", "Pending", {"cpu": 4, "memory": 1e9}, group="small"))
    s = b.flatten()
    tiers = PluginConf([[PluginOption("drf", enabled_preemptable=True, enabled_job_order=True)]])
    o, ev, order = kbo.cycle(s, tiers, actions=("preempt",), running=s.meta["running"])
    # s-new needs 4 cpu: ls = (1+4)/10 = 0.5; big after losing k tasks: 0.5, 0.4, ... -> only the FIRST preemptee (rs 0.5 >= ls 0.5) is a
    # victim: 1 cpu < 4 cpu -> validateVictims fails -> nothing happens
    assert o.result.evictions == 0 and int(o.decisions["kind"][0]) == 0
    # a 1-cpu preemptor: ls = 0.2; victims b0..b3 (rs 0.5, 0.4, 0.3, 0.2); one eviction is enough
    b.pods[-1] = B.Pod("ns", "s-new", "", "Pending", {"cpu": 1, "memory": 1e9}, group="small")
    s = b.flatten()
    o, ev, order = kbo.cycle(s, tiers, actions=("preempt",), running=s.meta["running"])
    assert o.result.evictions == 1 and int(o.decisions["kind"][0]) == 2
    # the victims queue pops the LOWEST TaskOrderFn task first: equal priority / creation -> highest UID among b0..b3
    names = s.meta["running"]["names"]
    assert {names[i] for i in np.nonzero(ev)[0]} == {"ns/b3"}
    np.testing.assert_allclose(o.job_share, [0.5, 0.2])


# ---- the remaining unit vectors of the reference that touch this path ----
def test_select_best_node_golden():
    """util/scheduler_helper_test.go:24-92: the pick must be ONE OF the max-score nodes (the reference picks randomly among them;
    the deterministic rule takes the first)."""
    import ctypes as C
    L = kbo.lib()
    L.kbo_select_best_node.restype = C.c_uint32
    for scores, expected in (([1.0, 1.0, 2.0, 2.0], {2, 3}), ([1.0, 1.0, 3.0, 2.0, 2.0], {2})):
        a = np.array(scores, dtype=np.float64)
        got = L.kbo_select_best_node(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint32(len(a)))
        assert got in expected and got == min(expected)


def test_arguments_get_int_golden():
    """framework/arguments_test.go:30-78: absent key and unparsable values keep the base value."""
    import ctypes as C
    L = kbo.lib()
    L.kbo_arguments_get_int.restype = C.c_int
    L.kbo_arguments_get_int.argtypes = [C.c_char_p, C.c_int]
    assert L.kbo_arguments_get_int(None, 10) == 10           # {"anotherkey": "12"}
    assert L.kbo_arguments_get_int(b"15", 10) == 15
    assert L.kbo_arguments_get_int(b"errorvalue", 11) == 11
    assert L.kbo_arguments_get_int(b"", 0) == 0
    # the engine's own parser (kb_build.h parse_int) must agree: an unparsable nodeorder weight keeps the default on both sides
    import util
    from kube_batch_b200 import synth
    s = synth.random_session(12, tasks=80, jobs=8, nodes=30)
    for bad in ("errorvalue", "", "3x", " 4"):
        conf = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"leastrequested.weight": bad, "mostrequested.weight": "2"}})
        o = kbo.allocate(s, conf)
        e = util.emu_allocate(s, conf, mode=1)
        util.assert_same_decisions(o.decisions, e.decisions, f"weight {bad!r}")


def test_job_info_add_task_golden():
    """api/job_info_test.go:35-101: p1 Pending, p2 Running on n1 (2000m/2G), p3 and p4 Pending WITH a node name = Bound (1000m/1G
    each) -> Allocated 4000m/4G, TaskStatusIndex {Pending: p1, Running: p2, Bound: p3, p4}."""
    b = B.SessionBuilder()
    b.add_node(B.build_node("n1", {"cpu": 8, "memory": 10e9}, pods=110))
    b.add_queue(B.Queue("q"))
    b.add_pod_group(B.PodGroup("c1", "uid", "q"))
    b.add_pod(B.Pod("c1", "p1", "", "Pending", {"cpu": 1, "memory": 1e9}, group="uid"))
    b.add_pod(B.Pod("c1", "p2", "n1", "Running", {"cpu": 2, "memory": 2e9}, group="uid"))
    b.add_pod(B.Pod("c1", "p3", "n1", "Pending", {"cpu": 1, "memory": 1e9}, group="uid"))
    b.add_pod(B.Pod("c1", "p4", "n1", "Pending", {"cpu": 1, "memory": 1e9}, group="uid"))
    s = b.flatten()
    assert s.T == 1 and s.meta["tasks"] == ["c1/p1"]
    assert s.job_alloc0[:2, 0].tolist() == [4000.0, 4e9] and s.job_ready0[0] == 3
    assert s.node_used[:2, 0].tolist() == [4000.0, 4e9] and s.node_pods[0] == 3


def test_new_resource_golden():
    """api/resource_info_test.go:27-57: cpu 4m -> MilliCPU 4, memory 2000 -> 2000, scalar quantities -> milli-units."""
    v, present = B.SessionBuilder._resource({"cpu": 0.004, "memory": 2000, "scalar.test/scalar1": 1, "hugepages-test": 2},
                                            ["cpu", "memory", "hugepages-test", "scalar.test/scalar1"])
    assert v.tolist() == [4.0, 2000.0, 2000.0, 1000.0] and present == 0b1100
    v, present = B.SessionBuilder._resource({}, ["cpu", "memory"])
    assert v.tolist() == [0.0, 0.0] and present == 0


def test_committed_cycle_digests_match_a_fresh_oracle_run():
    """tests/golden/cycle_hashes.json (what bench.py and the full-size GPU tests compare against) is reproducible: c2 here,
    c3 too (a few seconds); c4 is regenerated by make_golden.py only."""
    import json, os
    from kube_batch_b200 import digest, synth
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cycle_hashes.json")))
    for name in ("c2", "c3"):
        snap, conf = synth.make(name)
        o = kbo.allocate(snap, conf, threads=os.cpu_count() or 1)
        assert digest.decisions_digest(o.decisions) == g[name]["decisions"], name
        assert digest.state_digest(o.node_idle, o.node_releasing, o.job_ready, o.job_share) == g[name]["state"], name
        assert int(o.result.tasks_allocated) == g[name]["allocated"]


@pytest.mark.parametrize("replica", [2, 7])
def test_replica_digests_match_the_pipeline_emulation(replica):
    """bench.py --gpus N: rank r schedules synth.make("c3", replica=r) and checks itself against tests/golden/cycle_hashes.json
    "c3#r" (made by the oracle).  Here: the emulation of cycle_kernel's protocol (mode 5) reproduces those digests on the CPU."""
    import json
    import os
    import util
    from kube_batch_b200 import digest, synth
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cycle_hashes.json")))[f"c3#{replica}"]
    s, conf = synth.make("c3", replica=replica)
    e = util.emu_allocate(s, conf, mode=5)
    assert digest.decisions_digest(e.decisions) == g["decisions"]
    ns, os_ = util.emu_states(e)
    assert digest.state_digest(ns["idle"], ns["releasing"], os_["job_ready"], os_["job_share"]) == g["state"]
