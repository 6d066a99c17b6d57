"""CPU-only parity: the engine's host/device-shared logic (kb_core.h / kb_ctl.h / kb_build.h), re-enacted
step by step by tests/emu (scan -> top-K -> certified replay -> control -> gang commit), must reproduce
the oracle bit-exactly.  The CUDA thread mechanics themselves are covered by the `-m gpu` tests."""
import numpy as np
import pytest

from kube_batch_b200 import abi
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


def check(snap, conf, what, actions=1, mode=0):
    o = kbo.allocate(snap, conf, actions=actions)
    e = util.emu_allocate(snap, conf, actions=actions, mode=mode)
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


def test_session_reload_reuses_the_built_session():
    """kb_session_load again and again on one engine: the host-side BuiltSession (slabs, conf, tables) is recycled, nothing
    of the previous session may leak into the next — bigger, smaller, other plugins, other modes."""
    import ctypes as C
    L = util.emu_lib()
    L.kbemu_create2.restype = C.c_void_p
    L.kbemu_create2.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.kbemu_reload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.kbemu_run.argtypes = [C.c_void_p, C.c_uint32]
    L.kbemu_finish.argtypes = [C.c_void_p] + [C.c_void_p] * 14
    L.kbemu_destroy.argtypes = [C.c_void_p]
    seq = [(synth.random_session(1, tasks=300, jobs=30, nodes=200, queues=3), "default", 1, 1),
           (synth.random_session(2, tasks=20, jobs=3, nodes=7, queues=1), "none", 4, 1),
           (synth.random_session(3, tasks=500, jobs=60, nodes=400, queues=2, be_frac=0.3, be_variants=True), "c2", 2, 3),
           (synth.random_session(4, tasks=60, jobs=8, nodes=12, queues=4, R=5, W=3), "weights", 0, 1),
           (synth.random_session(5, tasks=250, jobs=25, nodes=90, queues=2), "allocate_test", 1, 3)]
    h = None
    try:
        for snap, cname, mode, actions in seq:
            conf = CONFS[cname]
            cs, k1 = snap.to_c()
            cc, k2 = conf.to_c()
            if h is None:
                h = L.kbemu_create2(C.addressof(cs), C.addressof(cc), 0, 1, mode)
                assert h
            else:
                assert L.kbemu_reload(h, C.addressof(cs), C.addressof(cc), mode) == 0
            assert L.kbemu_run(h, actions) == 0
            dec = np.zeros(max(snap.T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
            st = abi.kb_stats()
            L.kbemu_finish(h, dec.ctypes.data, C.addressof(st), *([None] * 12))
            o = kbo.allocate(snap, conf, actions=actions)
            util.assert_same_decisions(o.decisions, dec[:snap.T], f"reload {cname} mode{mode} actions{actions}")
    finally:
        if h:
            L.kbemu_destroy(h)


# ---------------- chained visits (visit_chain_kernel<K>): K classes per scan, look-ahead lists patched before their replay ----------------
@pytest.mark.parametrize("seed", range(12))
def test_chained_visits_random_sessions(seed):
    rng = np.random.default_rng(2000 + seed)
    tasks = int(rng.integers(5, 400))
    s = synth.random_session(seed + 700, tasks=tasks, jobs=int(rng.integers(1, min(tasks, 60) + 1)), nodes=int(rng.integers(1, 300)),
                             queues=int(rng.integers(1, 5)), min_member_frac=float(rng.choice([0.0, 0.5, 1.0])),
                             hetero=float(rng.choice([0, 0.3, 1.0])), prio_levels=int(rng.integers(1, 4)),
                             oversub=float(rng.choice([0.7, 1.3, 3.0])))
    for cname, conf in CONFS.items():
        for mode in (1, 2, 4):            # 1 = plain one-class launches (no overlap protocol)
            check(s, conf, f"chain seed{seed}/{cname}/K{mode}", mode=mode)


def test_chained_visits_save_launches_and_patch_small_clusters():
    # few nodes: every visit modifies nodes that sit in the look-ahead lists, so the patch path (drop + re-evaluate + floor) is hot
    s = synth.random_session(31, tasks=600, jobs=120, nodes=24, queues=1, min_member_frac=0.0, hetero=1.0, oversub=0.9)
    base = None
    for mode in (1, 2, 4):
        o, e = check(s, CONFS["default"], f"chain small/K{mode}", mode=mode)
        if mode == 1:
            base = e.result.kernel_launches
        else:
            assert e.result.chain_hits > 0 and e.result.kernel_launches < base
    s, conf = synth.make("c2")
    o, e1 = check(s, conf, "chain c2/K1", mode=1)
    o, e4 = check(s, conf, "chain c2/K4", mode=4)
    assert e4.result.kernel_launches * 3 < e1.result.kernel_launches       # single queue, static job order: predictions hit


# ---------------- persistent pipeline (cycle_kernel): stale look-ahead lists + patch, garbage keys for in-flight nodes ----------------
@pytest.mark.parametrize("seed", range(16))
def test_pipeline_protocol_random_sessions(seed):
    """kb_pipe.cuh: the list of a visit was scanned at an EARLIER log position (random lag <= 32 entries) from the scanners' own
    copy of the table; nodes modified since then carry arbitrary scanned keys (torn reads) and are dropped + re-evaluated."""
    rng = np.random.default_rng(4000 + seed)
    tasks = int(rng.integers(5, 400))
    s = synth.random_session(seed + 900, tasks=tasks, jobs=int(rng.integers(1, min(tasks, 60) + 1)), nodes=int(rng.integers(1, 300)),
                             queues=int(rng.integers(1, 5)), min_member_frac=float(rng.choice([0.0, 0.5, 1.0])),
                             hetero=float(rng.choice([0, 0.3, 1.0])), prio_levels=int(rng.integers(1, 4)),
                             oversub=float(rng.choice([0.7, 1.3, 3.0])))
    for cname, conf in CONFS.items():
        check(s, conf, f"pipe seed{seed}/{cname}", mode=5)
        check(s, conf, f"pipe seed{seed}/{cname}/+backfill", actions=3, mode=5)


def test_pipeline_protocol_small_clusters_and_baseline_configs():
    # few nodes: most of the list is in flight at every visit
    s = synth.random_session(31, tasks=600, jobs=120, nodes=24, queues=1, min_member_frac=0.0, hetero=1.0, oversub=0.9)
    check(s, CONFS["default"], "pipe small", mode=5)
    s = synth.random_session(33, tasks=900, jobs=90, nodes=40, queues=3, min_member_frac=0.5, hetero=0.3, oversub=1.1)
    check(s, CONFS["default"], "pipe small multi-queue", mode=5)
    for name in ("c1", "c2"):
        s, conf = synth.make(name)
        check(s, conf, "pipe " + name, mode=5)
    for (R, W) in [(8, 4), (5, 2)]:
        s = synth.random_session(60, tasks=150, jobs=15, nodes=300, queues=2, hetero=0.3, R=R, W=W)
        check(s, CONFS["default"], f"pipe R{R}W{W}", mode=5)


# ---------------- backfill (actions/backfill/backfill.go:40-71), the action after allocate in the default list ----------------
@pytest.mark.parametrize("seed", range(16))
def test_backfill_random_sessions(seed):
    rng = np.random.default_rng(1000 + seed)
    tasks = int(rng.integers(5, 300))
    s = synth.random_session(seed + 300, tasks=tasks, jobs=int(rng.integers(1, min(tasks, 40) + 1)), nodes=int(rng.integers(1, 200)),
                             queues=int(rng.integers(1, 5)), min_member_frac=float(rng.choice([0.0, 0.5, 1.0])),
                             hetero=float(rng.choice([0, 0.3, 1.0])), oversub=float(rng.choice([0.7, 1.3, 3.0])),
                             be_frac=float(rng.choice([0.1, 0.3, 0.9])), be_variants=True)
    for cname in ("default", "c2", "none", "allocate_test", "weights"):
        for actions in (2, 3):           # "backfill" alone and "allocate, backfill"
            check(s, CONFS[cname], f"backfill seed{seed}/{cname}/actions{actions}", actions=actions)


def test_backfill_first_feasible_node_pod_cap_and_gang():
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "a-gang3", "q", min_member=3))      # JobID order: ns/a-gang3 before ns/be
    b.add_pod_group(B.PodGroup("ns", "be", "q", min_member=1))
    b.add_node(B.Node("n0", {"cpu": 8, "memory": 32e9, "pods": 3}, labels={"zone": "a"}))
    b.add_node(B.Node("n1", {"cpu": 8, "memory": 32e9, "pods": 4}, labels={"zone": "b"}))
    req = {"cpu": 1, "memory": 1e9}
    # gang of 3: two regular pods (allocate places them, the job is NOT ready) + one best-effort pod (backfill completes it)
    b.add_pod(B.Pod("ns", "g-a", "", "Pending", req, group="a-gang3", creation=1))
    b.add_pod(B.Pod("ns", "g-b", "", "Pending", req, group="a-gang3", creation=2))
    b.add_pod(B.Pod("ns", "g-be", "", "Pending", {}, group="a-gang3", creation=3))
    # best-effort pods: first feasible node in name order, the pod cap moves them on, a selector nobody matches leaves one out
    for i in range(3):
        b.add_pod(B.Pod("ns", f"be-{i}", "", "Pending", {}, group="be", creation=10 + i))
    b.add_pod(B.Pod("ns", "be-zone-b", "", "Pending", {}, group="be", node_selector={"zone": "b"}, creation=20))
    b.add_pod(B.Pod("ns", "be-nowhere", "", "Pending", {}, group="be", node_selector={"zone": "z"}, creation=21))
    b.add_pod(B.Pod("ns", "be-late", "", "Pending", {}, group="be", creation=22))
    s = b.flatten()
    conf = PluginConf.default()
    o1, _ = check(s, conf, "backfill-hand/allocate", actions=1)
    o, e = check(s, conf, "backfill-hand/allocate+backfill", actions=3)
    name = {s.meta["tasks"][t]: t for t in range(s.T)}
    d1, d = o1.decisions, o.decisions
    node = lambda n: s.meta["nodes"][int(d["node"][name[n]])] if d["node"][name[n]] >= 0 else None
    # after allocate alone the gang is short of one member: nothing dispatched, the best-effort pod was skipped
    assert not d1["dispatched"][name["ns/g-a"]] and d1["kind"][name["ns/g-be"]] == abi.KB_KIND_SKIPPED
    assert {node("ns/g-a"), node("ns/g-b")} == {"n0", "n1"}              # least-requested spreads the two
    # backfill completes the gang: all three dispatched AT the backfill step of g-be
    for n in ("ns/g-a", "ns/g-b", "ns/g-be"):
        assert d["dispatched"][name[n]] and d["dispatch_step"][name[n]] == d["step"][name["ns/g-be"]]
    assert d["step"][name["ns/g-be"]] > max(d["step"][name["ns/g-a"]], d["step"][name["ns/g-b"]])
    got = [node(n) for n in ("ns/g-be", "ns/be-0", "ns/be-1", "ns/be-2", "ns/be-zone-b", "ns/be-nowhere", "ns/be-late")]
    # tasks go in UID order (be-0, be-1, be-2, be-late, be-nowhere, be-zone-b): be-late takes n1's last slot (cap 4)
    assert got == ["n0", "n0", "n1", "n1", None, None, "n1"]
    assert d["kind"][name["ns/be-nowhere"]] == abi.KB_KIND_NONE and d["kind"][name["ns/be-zone-b"]] == abi.KB_KIND_NONE
    assert o.node_pods.tolist() == [3, 4]


def test_over_committed_node_is_refused():
    # Idle < -epsilon cannot come out of the reference's cache (node.AddTask refuses what does not fit); the load says so
    s = synth.random_session(3, tasks=20, jobs=3, nodes=5)
    s.node_idle[0, 2] = -500.0
    s.invalidate()
    with pytest.raises(RuntimeError, match="over-committed"):
        util.emu_allocate(s, PluginConf.default())


def test_unknown_plugin_is_refused():
    s, _ = synth.make("c1")
    bad = PluginConf.from_names([["gang", "my-custom-plugin"]])
    with pytest.raises(RuntimeError):
        util.emu_allocate(s, bad)
    with pytest.raises(RuntimeError):
        kbo.allocate(s, bad)


def test_phantom_allocated_best_effort_task_matches_the_reference_order():
    """ssn.Allocate sets the task's status to Allocated BEFORE node.AddTask (framework/session.go:241-262); when AddTask refuses the
    task on every node that passed the predicates, the reference leaves a task that is Allocated but sits on no node (it counts
    towards JobReady and is dispatched by the job's next successful Allocate).  Only a best-effort task with a NON-ZERO request
    below the IsEmpty epsilons can get there (a truly empty Resreq always passes Resreq <= Idle because Idle > -epsilon).  Round 1
    pinned this as a divergence; the engine now carries a "some node passes ssn.PredicateFn" bit next to the candidate list."""
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "g", "q", min_member=1))
    b.add_node(B.Node("n0", {"cpu": 1, "memory": 4e9, "pods": 10}))
    b.add_pod(B.Pod("ns", "full", "n0", "Running", {"cpu": 1, "memory": 1e9}, group="g"))        # Idle cpu = 0
    b.add_pod(B.Pod("ns", "tiny-a", "", "Pending", {"cpu": 0.005}, group="g", creation=1))        # 5 m: IsEmpty, yet not zero
    b.add_pod(B.Pod("ns", "tiny-b", "", "Pending", {"cpu": 0.005}, group="g", creation=2))
    s = b.flatten()
    conf = PluginConf.from_names([["gang"], ["predicates"]])
    o = kbo.allocate(s, conf, actions=3)
    # tiny-a: 5 <= 0 within epsilon -> placed, Idle becomes -5.  tiny-b: |5 - (-5)| = 10 is not < 10 -> AddTask refuses it.
    assert o.decisions["kind"].tolist() == [abi.KB_KIND_ALLOCATED, abi.KB_KIND_ALLOCATED] and o.decisions["node"].tolist() == [0, -1]
    for mode in (0, 1, 5):
        check(s, conf, f"phantom mode{mode}", actions=3, mode=mode)
    # With the predicates plugin the phantom poisons the session: util.PodLister lists it, CachedNodeInfo.GetNodeInfo("") is an
    # error (plugins/util/util.go:93-100) that InterPodAffinityMatches returns for EVERY later pair (vendor/.../predicates.go:
    # 1381-1393, 1261-1270).  So (a) the task is lost on the FIRST node that passes the predicates — backfill.go:50-65 moves on to
    # the next node, whose predicate now fails — and (b) nobody is placed afterwards.
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "g", "q", min_member=1))
    b.add_node(B.Node("n0", {"cpu": 1, "memory": 4e9, "pods": 10}))
    b.add_node(B.Node("n1", {"cpu": 1, "memory": 4e9, "pods": 10}))                                # plenty of room, and never reached
    b.add_pod(B.Pod("ns", "full", "n0", "Running", {"cpu": 1, "memory": 1e9}, group="g"))
    for k, nm in enumerate(["a", "b", "c"]):
        b.add_pod(B.Pod("ns", "tiny-" + nm, "", "Pending", {"cpu": 0.005}, group="g", creation=k + 1))
    s = b.flatten()
    o = kbo.allocate(s, conf, actions=3)
    assert o.decisions["node"].tolist() == [0, -1, -1]
    assert o.decisions["kind"].tolist() == [abi.KB_KIND_ALLOCATED, abi.KB_KIND_ALLOCATED, abi.KB_KIND_NONE]
    for mode in (0, 1, 5):
        check(s, conf, f"poisoned session mode{mode}", actions=3, mode=mode)
    # without the predicates plugin nothing evaluates InterPodAffinityMatches: the loop simply tries the next node
    o = kbo.allocate(s, PluginConf.from_names([["gang"]]), actions=3)
    assert o.decisions["node"].tolist() == [0, 1, 1] and (o.decisions["kind"] == abi.KB_KIND_ALLOCATED).all()
    for mode in (0, 1, 5):
        check(s, PluginConf.from_names([["gang"]]), f"no predicates plugin mode{mode}", actions=3, mode=mode)
    # a phantom counts towards JobReady (gang), on one node and on several
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "g", "q", min_member=3))
    b.add_node(B.Node("n0", {"cpu": 1, "memory": 4e9, "pods": 10}))
    b.add_node(B.Node("n1", {"cpu": 1, "memory": 4e9, "pods": 1}))                                 # room for exactly one more pod
    b.add_pod(B.Pod("ns", "full", "n0", "Running", {"cpu": 1, "memory": 1e9}, group="g"))
    for k, nm in enumerate(["a", "b", "c", "d"]):
        b.add_pod(B.Pod("ns", "tiny-" + nm, "", "Pending", {"cpu": 0.005}, group="g", creation=k + 1))
    s = b.flatten()
    for mode in (0, 1, 5):
        check(s, conf, f"phantom gang mode{mode}", actions=3, mode=mode)


@pytest.mark.parametrize("seed", range(12))
def test_phantom_corner_random_clusters(seed):
    """Random small clusters full of sub-epsilon best-effort requests on exhausted nodes: the phantom corner at volume."""
    rng = np.random.default_rng(7000 + seed)
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    nn = int(rng.integers(1, 5))
    for n in range(nn):
        b.add_node(B.Node(f"n{n}", {"cpu": 1, "memory": 4e9, "pods": int(rng.choice([2, 3, 10]))}, labels={"zone": "ab"[n % 2]}))
    for g in range(int(rng.integers(1, 4))):
        b.add_pod_group(B.PodGroup("ns", f"g{g}", "q", min_member=int(rng.integers(0, 5))))
        if rng.random() < 0.8:
            b.add_pod(B.Pod("ns", f"g{g}-full", f"n{int(rng.integers(0, nn))}", "Running", {"cpu": float(rng.choice([0.99, 0.995, 1.0])), "memory": 1e9}, group=f"g{g}"))
        for k in range(int(rng.integers(1, 8))):
            req = {"cpu": float(rng.choice([0.0, 0.001, 0.005, 0.009]))}
            if req["cpu"] == 0.0:
                req = {}
            b.add_pod(B.Pod("ns", f"g{g}-p{k}", "", "Pending", req, group=f"g{g}", creation=k,
                            node_selector={"zone": str(rng.choice(["a", "b"]))} if rng.random() < 0.3 else {}))
    s = b.flatten()
    for conf in (PluginConf.from_names([["gang"], ["predicates"]]), PluginConf.default()):
        for actions in (2, 3):
            for mode in (1, 5):
                check(s, conf, f"phantom fuzz seed{seed}/actions{actions}/mode{mode}", actions=actions, mode=mode)


# ---------------- a12 NodeAffinityPriority: prototype of the engine algorithm (emulation only; the kernels follow next round) ----------------
def _pref_cluster(seed, pipe_geometry=False, nodes=None):
    """pipe_geometry: R = 3 (one scalar resource) and W = 2 atom words — the record geometry cycle_kernel is built for."""
    rng = np.random.default_rng(seed)
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    nn = int(rng.integers(2, 7)) if nodes is None else nodes
    zones = ["a", "b", "c"]
    for n in range(nn):
        alloc = {"cpu": float(rng.choice([2, 4, 8])), "memory": 64e9, "pods": int(rng.choice([3, 6, 110]))}
        if pipe_geometry:
            alloc["nvidia.com/gpu"] = 4
        b.add_node(B.Node(f"n{n:03d}", alloc, labels={"zone": zones[n % 3], "rank": str(n)}))
    for g in range(int(rng.integers(1, 4))):
        b.add_pod_group(B.PodGroup("ns", f"g{g}", "q", min_member=int(rng.integers(0, 3))))
        pref = [(int(rng.choice([0, 1, 20, 50, 100])), [("zone", "In", [str(rng.choice(zones))])]) for _ in range(int(rng.integers(1, 4)))]
        cpu = float(rng.choice([0.5, 1, 2]))
        for k in range(int(rng.integers(3, 14))):
            b.add_pod(B.Pod("ns", f"g{g}-p{k:02d}", "", "Pending", {"cpu": cpu, "memory": 1e9}, group=f"g{g}", creation=k,
                            preferred_terms=pref if rng.random() < 0.9 else []))
    return b.flatten(W=2 if pipe_geometry else 1)


@pytest.mark.parametrize("seed", range(40))
def test_preferred_node_affinity_two_pass_scan_prototype(seed):
    """The scan of a class with preferred terms runs in two passes (max count over the FEASIBLE nodes, then the keys with
    10*count/max added); the replay stops for a rescan as soon as the last feasible max-count node fills up, because every
    key of the launch used that normalisation.  Small clusters make that happen all the time."""
    s = _pref_cluster(6000 + seed)
    rescans = 0
    for conf in (PluginConf.default(),
                 PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "5"}}),
                 PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "-3"}}),
                 PluginConf.from_names([["gang", "priority"], ["predicates"]])):            # no nodeorder: the terms must not matter
        o, e = check(s, conf, f"pref seed{seed}", mode=1)
        rescans += e.result.rescans
    assert rescans >= 0


def test_a_placed_pod_with_inter_pod_affinity_terms_is_refused():
    """predicates.go:1261-1288: pods already on a node can reject it for OTHER pods through their anti-affinity terms.  The ABI
    carries that as a snapshot flag; the engine's host build (which the emulation runs) refuses such a session loudly."""
    s = synth.random_session(3)
    s.flags = abi.KB_SNAPSHOT_PLACED_POD_AFFINITY
    with pytest.raises(RuntimeError, match="placed pod carries inter-pod"):
        util.emu_allocate(s, PluginConf.default())


PREF_CONFS = (PluginConf.default(),
              PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "5"}}),
              PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "-3"}}),
              PluginConf.from_names([["gang", "priority"], ["predicates"]]))


@pytest.mark.parametrize("seed", range(40))
def test_preferred_node_affinity_in_the_pipeline_protocol(seed):
    """cycle_kernel's treatment (emulated): a class with preferred terms only ever uses a list of the current table state;
    the scanners run pass 1 (max count over the feasible nodes), exchange, then build the keys; the replayer counts the
    feasible max-count nodes down."""
    s = _pref_cluster(6100 + seed, pipe_geometry=True, nodes=(None if seed % 4 else 150))
    assert s.R == 3 and s.W == 2
    for conf in PREF_CONFS:
        check(s, conf, f"pref pipe seed{seed}", mode=5)
        check(s, conf, f"pref pipe seed{seed} plain", mode=1)


def test_preferred_node_affinity_normalisation_goes_stale_exactly_when_the_last_max_node_fills():
    # n-a (count 100) takes two pods (pod cap 2).  n-b (count 20) is 40 % busy, n-c (count 0) is empty, so n-c's resource scores
    # beat n-b's by 4..6 points: while n-a is feasible b's affinity term is 10*20/100 = 2 and n-c would win; once n-a is full the
    # max count is 20, b's term jumps to 10 and n-b wins.  A replay that kept the stale normalisation would pick n-c.
    b = B.SessionBuilder()
    b.add_queue(B.Queue("q", 1))
    b.add_pod_group(B.PodGroup("ns", "g", "q", min_member=5))       # not ready before the 5th pod: the whole job is ONE run of one launch
    b.add_pod_group(B.PodGroup("ns", "old", "q", min_member=1))
    b.add_node(B.Node("n-a", {"cpu": 64, "memory": 256e9, "pods": 2}, labels={"zone": "a"}))
    b.add_node(B.Node("n-b", {"cpu": 64, "memory": 256e9, "pods": 110}, labels={"zone": "b"}))
    b.add_node(B.Node("n-c", {"cpu": 64, "memory": 256e9, "pods": 110}, labels={"zone": "c"}))
    b.add_pod(B.Pod("ns", "busy", "n-b", "Running", {"cpu": 25, "memory": 1e9}, group="old"))
    pref = [(100, [("zone", "In", ["a"])]), (20, [("zone", "In", ["b"])])]
    for k in range(5):
        b.add_pod(B.Pod("ns", f"p{k}", "", "Pending", {"cpu": 1, "memory": 1e9}, group="g", creation=k, preferred_terms=pref))
    s = b.flatten()
    conf = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]])
    fit, score = kbo.predicate_score(s, conf, 0)
    assert score[2] - score[1] > 2 and score[2] - score[1] < 10 and score[0] > score[2]      # the fixture is in the sensitive band
    o, e = check(s, conf, "pref stale", mode=1)
    names = [s.meta["nodes"][n] for n in o.decisions["node"]]
    assert names == ["n-a", "n-a", "n-b", "n-b", "n-b"]
    assert e.result.rescans >= 1


# ---------------- exact-arithmetic shortcuts of kb_core.h ----------------
def test_le_func_single_subtraction_equals_the_reference_form():
    """kb_core.h le_func evaluates `l < r || |l - r| < eps` (api/resource_info.go:268-274) as `(l - r) < eps`.  Same predicate on
    the reference's LessEqual vectors (resource_info_test.go:246-304), around every epsilon boundary, and on random values."""
    import ctypes as C
    L = util.emu_lib()
    for f in (L.kbemu_le, L.kbemu_le_reference_form):
        f.argtypes = [C.c_double, C.c_double, C.c_double]
        f.restype = C.c_int
    eps = [10.0, 10.0 * 1024 * 1024]
    vals = [0.0, -0.0, 1.0, 4.0, 10.0, 2000.0, 4000.0, 8000.0, 1e9, 4e9, 10.0 * 1024 * 1024, 1e300, -1e300, 5e-324, -5e-324, 9.999999999999998, 10.000000000000002]
    rng = np.random.default_rng(3)
    cases = [(l, r, e) for l in vals for r in vals for e in eps]
    for e in eps:
        for base in (0.0, 1000.0, 3e9, 1e15):
            for d in (e, np.nextafter(e, 0), np.nextafter(e, 1e300), -e, e / 2, 0.0):
                cases.append((base + d, base, e)); cases.append((base, base + d, e))
    cases += [(float(a), float(b), float(e)) for a, b, e in zip(rng.integers(0, 1 << 40, 20000), rng.integers(0, 1 << 40, 20000), rng.choice(eps, 20000))]
    cases += [(float(a), float(a + d), 10.0) for a, d in zip(rng.integers(0, 100000, 20000), rng.integers(-12, 13, 20000))]
    for l, r, e in cases:
        assert L.kbemu_le(l, r, e) == L.kbemu_le_reference_form(l, r, e), (l, r, e)


def test_div_0_to_10_estimate_plus_fixup_is_exact():
    """least / most requested use (x * 10) / capacity in Go int64 arithmetic (least_requested.go:49-58): the engine's
    float-estimate + one fix-up each way must equal the integer quotient, including at every exact multiple and its neighbours."""
    import ctypes as C
    L = util.emu_lib()
    L.kbemu_div_0_to_10.argtypes = [C.c_longlong, C.c_longlong]
    L.kbemu_div_0_to_10.restype = C.c_longlong
    rng = np.random.default_rng(5)
    bs = [1, 2, 3, 7, 10, 999, 1000, 32000, 96000, 128 << 30, 384 << 30, (1 << 58) - 1, 1 << 58] + [int(x) for x in rng.integers(1, 1 << 58, 3000)]
    for b in bs:
        for k in range(0, 11):
            for d in (-2, -1, 0, 1, 2):
                a = k * b + d
                if 0 <= a <= 10 * b:
                    assert L.kbemu_div_0_to_10(a, b) == a // b, (a, b)
        for a in rng.integers(0, 10 * b + 1, 20, dtype=np.uint64 if 10 * b >= 1 << 63 else np.int64):
            assert L.kbemu_div_0_to_10(int(a), b) == int(a) // b, (int(a), b)
