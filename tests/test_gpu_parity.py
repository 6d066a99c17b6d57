"""`-m gpu`: the CUDA path, called through the C ABI (libkbgpu.so via ctypes), against the CPU oracle,
the committed golden fixtures and size-independent properties.  Bit-exact: every field of every decision."""
import json
import os

import numpy as np
import pytest

from kube_batch_b200 import abi, engine, synth
from kube_batch_b200.snapshot import PluginConf
from oracle import kbo
import cases
import util
from test_emu_parity import CONFS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(device=0)
    yield e
    e.close()


def run_and_check(eng, snap, conf, what, states=True):
    o = kbo.allocate(snap, conf)
    eng.load(snap, conf)
    r = eng.allocate()
    util.assert_same_decisions(o.decisions, r.decisions, what)
    if states:
        util.assert_same_state(o, eng.node_state(), eng.order_state(), what)
    st = r.stats
    assert st.tasks_processed == o.result.tasks_processed
    assert st.tasks_allocated == o.result.tasks_allocated
    assert st.tasks_pipelined == o.result.tasks_pipelined
    assert st.visits == o.result.visits
    assert st.jobs_ready == o.result.jobs_ready
    assert st.pairs_logical == o.result.pairs_logical
    return o, r


def test_reference_allocate_test_cases(eng):
    exp = json.load(open(os.path.join(GOLD, "allocate_test_expected.json")))
    for name, mk in (("case1", cases.allocate_test_case1), ("case2", cases.allocate_test_case2)):
        s = mk()
        eng.load(s, cases.tiers_allocate_test())
        r = eng.allocate()
        binds = {s.meta["tasks"][t]: s.meta["nodes"][n] for t, n in r.bind_map().items()}
        assert binds == exp[name], name


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_baseline_configs_vs_oracle_and_golden(eng, name):
    s, conf = synth.make(name)
    o, r = run_and_check(eng, s, conf, name)
    g = np.load(os.path.join(GOLD, f"{name}_oracle.npz"))
    util.assert_same_decisions(g["decisions"], r.decisions, name + " (committed golden)")
    np.testing.assert_array_equal(g["node_idle"], eng.node_state()["idle"])


def test_allocate_is_repeatable(eng):
    s, conf = synth.make("c2")
    eng.load(s, conf)
    a = eng.allocate()
    b = eng.allocate()
    util.assert_same_decisions(a.decisions, b.decisions, "second kb_allocate on the same session")


@pytest.mark.parametrize("seed", range(16))
def test_random_sessions_all_confs(eng, seed):
    rng = np.random.default_rng(seed)
    tasks = int(rng.integers(5, 300))
    jobs = int(rng.integers(1, min(tasks, 40) + 1))
    s = synth.random_session(seed, tasks=tasks, jobs=jobs, nodes=int(rng.integers(1, 200)), queues=int(rng.integers(1, 5)),
                             min_member_frac=float(rng.choice([0.0, 0.5, 1.0])), hetero=float(rng.choice([0, 0.3, 1.0])),
                             prio_levels=int(rng.integers(1, 4)), oversub=float(rng.choice([0.7, 1.3, 3.0])))
    for cname, conf in CONFS.items():
        run_and_check(eng, s, conf, f"seed{seed}/{cname}")


def test_overlap_mode_is_bit_exact():
    """KB_ENGINE_FORCE_OVERLAP: scanners run ahead on the predicted class while one CTA replays (exclusion + patch)."""
    e = engine.Engine(device=0, flags=2)
    try:
        for name in ("c2",):
            s, conf = synth.make(name)
            run_and_check(e, s, conf, name + "/overlap")
        for seed in range(6):
            s = synth.random_session(100 + seed, tasks=200 + 40 * seed, jobs=12 + seed, nodes=150 + 60 * seed, queues=1 + seed % 3,
                                     min_member_frac=[0.0, 0.5, 1.0][seed % 3], hetero=[0, 0.3, 1.0][seed % 3], prio_levels=2)
            for cname in ("default", "c2", "nogang", "weights"):
                run_and_check(e, s, CONFS[cname], f"overlap seed{seed}/{cname}")
        s = synth.generate(synth.SynthSpec("wide", tasks=600, jobs=60, nodes=148 * 128 * 2 + 77, seed=99))
        run_and_check(e, s, PluginConf.default(), "wide/overlap")
    finally:
        e.close()


@pytest.mark.parametrize("flag", [abi.KB_ENGINE_CHAIN2, abi.KB_ENGINE_CHAIN4])
def test_chained_visits_are_bit_exact(flag):
    """visit_chain_kernel<2|4>: K classes per scan; following visits replayed from patched look-ahead lists."""
    e = engine.Engine(device=0, flags=flag)
    try:
        s, conf = synth.make("c2")
        o, r = run_and_check(e, s, conf, "c2/chain")
        assert r.stats.chain_hits > 0
        for seed in range(6):
            s = synth.random_session(100 + seed, tasks=200 + 40 * seed, jobs=12 + seed, nodes=150 + 60 * seed, queues=1 + seed % 3,
                                     min_member_frac=[0.0, 0.5, 1.0][seed % 3], hetero=[0, 0.3, 1.0][seed % 3], prio_levels=2)
            for cname in ("default", "c2", "nogang", "weights"):
                run_and_check(e, s, CONFS[cname], f"chain seed{seed}/{cname}")
        s = synth.random_session(31, tasks=600, jobs=120, nodes=24, queues=1, min_member_frac=0.0, hetero=1.0, oversub=0.9)
        run_and_check(e, s, CONFS["default"], "chain small cluster (patch path)")
        s = synth.generate(synth.SynthSpec("wide", tasks=600, jobs=60, nodes=148 * 128 * 2 + 77, seed=99))
        run_and_check(e, s, PluginConf.default(), "wide/chain")
        for (R, W) in [(8, 4), (5, 2)]:
            s = synth.random_session(60, tasks=150, jobs=15, nodes=300, queues=2, hetero=0.3, R=R, W=W)
            run_and_check(e, s, CONFS["default"], f"chain R{R}W{W}")
        s, conf = synth.make("c3")
        run_and_check(e, s, conf, "c3/chain")
    finally:
        e.close()


@pytest.mark.parametrize("R,W", [(4, 3), (6, 4), (8, 4), (5, 2)])
def test_wide_records_more_dims_and_mask_words(eng, R, W):
    """Generic tile geometry: ncols = 2R + 6 + 3W up to 34 columns -> wider TMA tiles, fewer tiles per scan iteration."""
    for seed in range(3):
        s = synth.random_session(seed + 50, tasks=150, jobs=15, nodes=300 + 700 * seed, queues=2, hetero=0.3, R=R, W=W)
        for cname in ("default", "c2"):
            run_and_check(eng, s, CONFS[cname], f"R{R}W{W}/seed{seed}/{cname}")
    eng.load(s, CONFS["c2"])              # back to the snapshot's initial node state
    fit, score = eng.predicate_score(0, min(s.T, 40))
    for t in range(0, min(s.T, 40), 7):
        of, osc = kbo.predicate_score(s, CONFS["c2"], t)
        np.testing.assert_array_equal(of, fit[t])
        np.testing.assert_array_equal(osc, score[t])


def test_long_run_forces_rescans(eng):
    s = synth.random_session(7, tasks=400, jobs=1, nodes=300, hetero=0.0, oversub=0.5)
    run_and_check(eng, s, synth.conf_c2(), "long-run")


def test_multi_tile_grid(eng):
    # > 148 tiles: every scan CTA walks several TMA tiles (double buffer) and the merge sees 148 lists
    s = synth.generate(synth.SynthSpec("wide", tasks=600, jobs=60, nodes=148 * 128 * 2 + 77, seed=99))
    run_and_check(eng, s, PluginConf.default(), "wide", states=True)


def test_predicate_score_matrix_vs_oracle(eng):
    s, conf = synth.make("c2")
    eng.load(s, conf)
    fit, score = eng.predicate_score(0, s.T)
    for t in list(range(0, s.T, 97)) + [s.T - 1]:
        of, osc = kbo.predicate_score(s, conf, t)
        np.testing.assert_array_equal(of, fit[t], err_msg=f"fit row {t}")
        np.testing.assert_array_equal(osc, score[t], err_msg=f"score row {t}")
    # K3: the fused argmax equals first-max of the matrix row (SelectBestNode with the deterministic rule)
    best = engine.key_node(eng.best_nodes(0, s.T))
    masked = np.where(fit > 0, score, -1.0)
    exp = np.where(fit.any(axis=1), masked.argmax(axis=1), -1)
    np.testing.assert_array_equal(exp, best)


def test_degenerate_sessions(eng):
    from kube_batch_b200.snapshot import Snapshot
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
        run_and_check(eng, s, PluginConf.default(), f"degenerate T{T} J{J} N{N}")


def test_unknown_plugin_and_unsupported_feature_fail_loudly(eng):
    s, _ = synth.make("c1")
    with pytest.raises(engine.KbError) as ei:
        eng.load(s, PluginConf.from_names([["gang", "my-custom-plugin"]]))
    assert ei.value.code == abi.KB_E_UNSUPPORTED_PLUGIN
    s.task_flags[0] |= abi.KB_TASK_HAS_POD_AFFINITY
    with pytest.raises(engine.KbError) as ei:
        eng.load(s, PluginConf.default())
    assert ei.value.code == abi.KB_E_UNSUPPORTED_FEATURE


# ---------------- bind fan-out list (framework/session.go:277-314 -> cache.Bind) ----------------
@pytest.mark.parametrize("name", ["c2", "c3"])
def test_bind_list_is_the_dispatch_order_of_the_oracle(eng, name):
    s, conf = synth.make(name)
    o = kbo.allocate(s, conf)
    eng.load(s, conf)
    eng.allocate()
    task, node = eng.bind_list()
    d = o.decisions
    disp = np.nonzero(d["dispatched"] & (d["node"] >= 0))[0]
    order = disp[np.lexsort((d["step"][disp], d["dispatch_step"][disp]))]          # by dispatch step, then Allocate order
    np.testing.assert_array_equal(task, order.astype(np.uint32))
    np.testing.assert_array_equal(node, d["node"][order])
    assert len(task) == int(d["dispatched"].sum()) > 0


# ---------------- a12 NodeAffinityPriority (node_affinity.go:34-77 + reduce.go:28-63) in cycle_kernel ----------------
@pytest.mark.parametrize("seed", range(12))
def test_preferred_node_affinity_on_the_gpu(eng, seed):
    """Preferred node-affinity terms: the scanner CTAs find the max count over the feasible nodes (pass 1 + one exchange between
    the CTAs), build the keys with weight * (10 * count / max), and the replayer counts the feasible max-count nodes down."""
    from test_emu_parity import _pref_cluster, PREF_CONFS
    s = _pref_cluster(6100 + seed, pipe_geometry=True, nodes=(None if seed % 3 else 400))
    for conf in PREF_CONFS:
        run_and_check(eng, s, conf, f"pref seed{seed}")


def test_preferred_node_affinity_outside_the_pipeline_geometry(eng):
    """R = 2, W = 1: the per-visit kernels run it (visit_kernel<0,1> + aff_prepass_kernel: max count over the feasible nodes before the
    visit, a fresh scan per task of a class with preferred terms)."""
    from test_emu_parity import _pref_cluster, PREF_CONFS
    for seed in range(6):
        s = _pref_cluster(6000 + seed)
        assert s.R == 2
        for conf in PREF_CONFS:
            o, r = run_and_check(eng, s, conf, f"pref outside the pipeline geometry seed{seed}")
            assert r.stats.pipeline == 0


# ---------------- kb_backfill (actions/backfill/backfill.go:40-71) ----------------
def run_backfill_and_check(eng, snap, conf, what, actions):
    o = kbo.allocate(snap, conf, actions=actions)
    eng.load(snap, conf)
    r = eng.allocate() if actions & 1 else None
    r = eng.backfill()
    util.assert_same_decisions(o.decisions, r.decisions, what)
    util.assert_same_state(o, eng.node_state(), eng.order_state(), what)
    st = r.stats
    assert (st.tasks_processed, st.tasks_allocated, st.tasks_pipelined, st.visits, st.jobs_ready, st.pairs_logical) == \
        (o.result.tasks_processed, o.result.tasks_allocated, o.result.tasks_pipelined, o.result.visits, o.result.jobs_ready,
         o.result.pairs_logical), what
    return o, r


@pytest.mark.parametrize("seed", range(8))
def test_backfill_random_sessions(eng, seed):
    rng = np.random.default_rng(1000 + seed)
    tasks = int(rng.integers(5, 300))
    s = synth.random_session(seed + 300, tasks=tasks, jobs=int(rng.integers(1, min(tasks, 40) + 1)), nodes=int(rng.integers(1, 200)),
                             queues=int(rng.integers(1, 5)), min_member_frac=float(rng.choice([0.0, 0.5, 1.0])),
                             hetero=float(rng.choice([0, 0.3, 1.0])), oversub=float(rng.choice([0.7, 1.3, 3.0])),
                             be_frac=float(rng.choice([0.1, 0.3, 0.9])), be_variants=True)
    for cname in ("default", "c2", "allocate_test"):
        for actions in (2, 3):
            run_backfill_and_check(eng, s, CONFS[cname], f"backfill seed{seed}/{cname}/actions{actions}", actions)


def test_backfill_phantom_allocated_tasks_on_the_gpu(eng):
    """ssn.Allocate's status-before-AddTask order (session.go:241-262): sub-epsilon best-effort requests on exhausted nodes leave
    tasks Allocated on no node; same random clusters as tests/test_emu_parity.py::test_phantom_corner_random_clusters."""
    from kube_batch_b200 import builder as B
    phantoms = 0
    for seed in range(12):
        rng = np.random.default_rng(7000 + seed)
        b = B.SessionBuilder()
        b.add_queue(B.Queue("q", 1))
        nn = int(rng.integers(1, 5))
        for n in range(nn):
            b.add_node(B.Node(f"n{n}", {"cpu": 1, "memory": 4e9, "pods": int(rng.choice([2, 3, 10]))}, labels={"zone": "ab"[n % 2]}))
        for g in range(int(rng.integers(1, 4))):
            b.add_pod_group(B.PodGroup("ns", f"g{g}", "q", min_member=int(rng.integers(0, 5))))
            if rng.random() < 0.8:
                b.add_pod(B.Pod("ns", f"g{g}-full", f"n{int(rng.integers(0, nn))}", "Running",
                                {"cpu": float(rng.choice([0.99, 0.995, 1.0])), "memory": 1e9}, group=f"g{g}"))
            for k in range(int(rng.integers(1, 8))):
                req = {"cpu": float(rng.choice([0.0, 0.001, 0.005, 0.009]))}
                if req["cpu"] == 0.0:
                    req = {}
                b.add_pod(B.Pod("ns", f"g{g}-p{k}", "", "Pending", req, group=f"g{g}", creation=k,
                                node_selector={"zone": str(rng.choice(["a", "b"]))} if rng.random() < 0.3 else {}))
        s = b.flatten()
        for conf in (PluginConf.from_names([["gang"], ["predicates"]]), PluginConf.default()):
            for actions in (2, 3):
                o, r = run_backfill_and_check(eng, s, conf, f"phantom seed{seed}/actions{actions}", actions)
                phantoms += int(((r.decisions["kind"] == abi.KB_KIND_ALLOCATED) & (r.decisions["node"] == -1)).sum())
    assert phantoms > 0


def test_backfill_multi_tile_then_allocate_restarts(eng):
    # many nodes (several tiles), many best-effort pods; afterwards kb_allocate must restart from the loaded state
    s = synth.random_session(77, tasks=1500, jobs=60, nodes=3000, queues=2, min_member_frac=0.5, hetero=0.3, oversub=1.3,
                             be_frac=0.3, be_variants=True)
    conf = PluginConf.default()
    o, r = run_backfill_and_check(eng, s, conf, "backfill multi-tile", 3)
    again = eng.backfill()                       # nothing left to do: same table
    util.assert_same_decisions(r.decisions, again.decisions, "second kb_backfill")
    a = eng.allocate()
    util.assert_same_decisions(kbo.allocate(s, conf).decisions, a.decisions, "kb_allocate after kb_backfill")


def test_c3_full_size_parity_and_properties(eng):
    """BASELINE config 3 (50k tasks x 5k nodes, default tiers) — full-size parity and invariants."""
    s, conf = synth.make("c3")
    o, r = run_and_check(eng, s, conf, "c3", states=True)
    d = r.decisions
    ns = eng.node_state()
    # no node over-committed: Idle never below -epsilon, Used + Idle == Allocatable for untouched releasing
    assert (ns["idle"][0] > -10).all() and (ns["idle"][1] > -10 * 1024 * 1024).all()
    # gang: a job's tasks are dispatched iff the job reached MinAvailable
    tj = s.job_of_task()
    osr = eng.order_state()
    ready = osr["job_ready"] >= s.job_min_avail
    alloc = d["kind"] == abi.KB_KIND_ALLOCATED
    assert (d["dispatched"][alloc] == ready[tj[alloc]]).all()
    assert not d["dispatched"][~alloc].any()
    # steps are a permutation of 0..placed-1
    placed = d["step"] != 0xFFFFFFFF
    assert sorted(d["step"][placed].tolist()) == list(range(int(placed.sum())))
    assert r.stats.pairs_logical == int(r.stats.tasks_processed) * s.N
    check_committed_digest(eng, "c3", r)


def check_committed_digest(eng, name, r):
    """The committed oracle digests (tests/golden/cycle_hashes.json, made HERE by make_golden.py) — what bench.py checks at every N."""
    from kube_batch_b200 import digest
    g = json.load(open(os.path.join(GOLD, "cycle_hashes.json")))[name]
    ns, osr = eng.node_state(), eng.order_state()
    assert digest.decisions_digest(r.decisions) == g["decisions"], name + ": decisions differ from the committed oracle digest"
    assert digest.state_digest(ns["idle"], ns["releasing"], osr["job_ready"], osr["job_share"]) == g["state"], name + ": state digest"


def test_c5_shaped_session_at_one_fiftieth_scale(eng):
    """BASELINE config 5 (1M tasks x 100k nodes, one node shape) is too large for the CPU oracle; its SHAPE is not: the same
    generator at 1/50 scale (20k tasks / 2k PodGroups / 2k homogeneous nodes, gang + drf + predicates + nodeorder) with full
    oracle parity, and at 12 500 nodes (several resident tiles per scanner CTA... 98 tiles on 98 SMs) as well."""
    for scale, nodes in ((50, 2_000), (8, 12_500)):
        spec = synth.SynthSpec(f"c5/{scale}", tasks=1_000_000 // scale, jobs=100_000 // scale, nodes=nodes, homogeneous_nodes=True, seed=0xB200 + 5)
        s = synth.generate(spec)
        conf = synth.config_conf("c5")
        run_and_check(eng, s, conf, f"c5 shape 1/{scale}")


def test_c4_multiqueue_full_size_parity(eng):
    """BASELINE config 4 (200k tasks x 20k nodes, 8 queues with proportion): exact Go-heap replay with stale keys at scale.
    The oracle runs here as well (about a minute on 16 host threads)."""
    s, conf = synth.make("c4")
    o = kbo.allocate(s, conf, threads=min(16, os.cpu_count() or 1))
    eng.load(s, conf)
    r = eng.allocate()
    util.assert_same_decisions(o.decisions, r.decisions, "c4")
    util.assert_same_state(o, eng.node_state(), eng.order_state(), "c4")
    check_committed_digest(eng, "c4", r)
    print(f"c4: gpu {r.stats.gpu_ms:.1f} ms, oracle {o.result.seconds:.1f} s, scans {r.stats.scans}, visits {r.stats.visits}")
