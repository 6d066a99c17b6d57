"""`-m gpu`: inter-pod (anti)affinity on the CUDA path (visit_kernel<BF, AFF = 1> + aff_prepass_kernel) through the C ABI,
bit-exact against the oracle, which walks the raw pod objects (predicates.go:1261-1572, interpod_affinity.go:99-235)."""
import numpy as np
import pytest

from kube_batch_b200 import abi, builder as B, engine
from kube_batch_b200.snapshot import PluginConf
from oracle import kbo
import aff_gen
import util
import test_pod_affinity as tpa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(device=0)
    yield e
    e.close()


def run_and_check(eng, snap, conf, what, actions=1):
    o = kbo.allocate(snap, conf, actions=actions)
    eng.load(snap, conf)
    r = eng.allocate() if actions & 1 else None
    if actions & 2:
        r = eng.backfill()
    util.assert_same_decisions(o.decisions, r.decisions, what)
    util.assert_same_state(o, eng.node_state(), eng.order_state(), what)
    st = r.stats
    assert (st.tasks_processed, st.tasks_allocated, st.tasks_pipelined, st.jobs_ready, st.pairs_logical) == \
        (o.result.tasks_processed, o.result.tasks_allocated, o.result.tasks_pipelined, o.result.jobs_ready, o.result.pairs_logical), what
    assert st.pipeline == 0 and st.kernel_launches > 0
    return o, r


def test_hand_computed_cases_on_the_gpu(eng):
    # self anti-affinity: one pod per host
    sb = tpa.cluster(4)
    for i in range(5):
        p = tpa.pod(f"p{i}", {"app": "web"}, creation=i)
        p.pod_anti_affinity = B.PodAffinity(required=[tpa.term(tpa.HOST, app="web")])
        sb.add_pod(p)
    o, r = run_and_check(eng, sb.flatten(), PluginConf.default(), "spread")
    assert sorted(r.decisions["node"][:4].tolist()) == [0, 1, 2, 3] and r.decisions["kind"][4] == abi.KB_KIND_NONE
    # first pod of a self-affine series, then the zone is fixed
    sb = tpa.cluster(6, zones=3)
    for i in range(4):
        p = tpa.pod(f"p{i}", {"app": "ring"}, creation=i)
        p.pod_affinity = B.PodAffinity(required=[tpa.term(tpa.ZONE, app="ring")])
        sb.add_pod(p)
    o, r = run_and_check(eng, sb.flatten(), PluginConf.default(), "ring")
    assert len({int(n) % 3 for n in r.decisions["node"]}) == 1
    # the GetNodeInfo quirk of the priority (nodeorder.go:49-63)
    sb = tpa.cluster(3, zones=3)
    sb.add_pod_group(B.PodGroup("ns", "pg0", "q1", min_member=1, creation=0))
    sb.add_pod_group(B.PodGroup("ns", "pg2", "q1", min_member=1, creation=2))
    sb.pod_groups[0].creation = 1
    c0 = tpa.pod("c0", {"app": "c"}, group="pg0"); c0.node_selector = {tpa.HOST: "n0"}
    a0 = tpa.pod("a0", {"app": "a"}, group="pg1"); a0.node_selector = {tpa.HOST: "n2"}
    b0 = tpa.pod("b0", {"app": "b"}, group="pg2"); b0.pod_affinity = B.PodAffinity(preferred=[(7, tpa.term(tpa.HOST, app="a"))])
    for p in (c0, a0, b0):
        sb.add_pod(p)
    snap = sb.flatten()
    o, r = run_and_check(eng, snap, tpa.ONLY_PODAFF, "first unbound node")
    where = {snap.meta["tasks"][t]: int(r.decisions["node"][t]) for t in range(snap.T)}
    assert where == {"ns/c0": 0, "ns/a0": 2, "ns/b0": 0}


@pytest.mark.parametrize("seed", range(32))
def test_random_affinity_sessions_on_the_gpu(eng, seed):
    sb = aff_gen.random_affinity_session(seed, n_nodes=4 + seed % 13, n_groups=3 + seed % 6, besteffort=seed % 4 == 3)
    snap = sb.flatten()
    if snap.pod_affinity is None:
        pytest.skip("no affinity terms drawn")
    for ci, conf in enumerate(tpa.AFF_CONFS):
        run_and_check(eng, snap, conf, f"seed {seed} conf {ci}", actions=1)
        run_and_check(eng, snap, conf, f"seed {seed} conf {ci} +backfill", actions=3)


@pytest.mark.parametrize("seed", range(4))
def test_larger_affinity_sessions_on_the_gpu(eng, seed):
    """several node tiles per scan CTA group, hundreds of per-task scans"""
    sb = aff_gen.random_affinity_session(9000 + seed, n_nodes=700 + 97 * seed, n_groups=60, p_affine=0.5, spec_pool=5)
    snap = sb.flatten()
    assert snap.pod_affinity is not None
    o, r = run_and_check(eng, snap, PluginConf.default(), f"large seed {seed}", actions=3)
    assert int((r.decisions["kind"] == abi.KB_KIND_ALLOCATED).sum()) > 50


def test_cycle_with_the_placing_actions_and_refusals(eng):
    sb = aff_gen.random_affinity_session(77, n_nodes=20, n_groups=10)
    snap = sb.flatten()
    assert snap.pod_affinity is not None
    conf = PluginConf.default()
    o = kbo.allocate(snap, conf, actions=3)
    eng.load(snap, conf)
    r = eng.cycle(("allocate", "backfill"))
    res = r[0] if isinstance(r, tuple) else r
    util.assert_same_decisions(o.decisions, res.decisions, "kb_cycle(allocate, backfill)")
    # reclaim / preempt do not maintain the affinity counters: refused loudly
    with pytest.raises(engine.KbError) as ei:
        eng.load_running(snap.meta["running"])
    assert ei.value.code == abi.KB_E_UNSUPPORTED_FEATURE
    with pytest.raises(engine.KbError) as ei:
        eng.predicate_score(0, 1)                  # the priority terms need reductions: only `fit` is offered for such sessions
    assert ei.value.code == abi.KB_E_UNSUPPORTED_FEATURE
    # the flags alone, without the flattened tables: refused
    snap.pod_affinity = None
    with pytest.raises(engine.KbError) as ei:
        eng.load(snap, conf)
    assert ei.value.code == abi.KB_E_UNSUPPORTED_FEATURE


def test_big_affinity_session_matches_the_emulation_and_is_timed(eng):
    """2 000 nodes / ~3 000 pending pods, a third of the PodGroups with (anti)affinity terms: hundreds of fresh scans, every one with
    predicate step 10, many with the three priority passes.  The object-level oracle is O(pods) per pair (the reference's cost), so at
    this size the engine is compared with the emulation of the device algorithm (itself pinned on the oracle at small sizes)."""
    import json, os, time
    sb = aff_gen.random_affinity_session(4242, n_nodes=2000, n_groups=1000, p_affine=0.35, spec_pool=6)
    snap = sb.flatten()
    assert snap.pod_affinity is not None
    conf = PluginConf.default()
    e = util.emu_allocate(snap, conf, actions=1, mode=1)
    eng.load(snap, conf)
    eng.allocate()
    t0 = time.perf_counter()
    r = eng.allocate()
    wall = time.perf_counter() - t0
    util.assert_same_decisions(e.decisions, r.decisions, "big affinity session vs emulation")
    st = r.stats
    rec = {"nodes": int(snap.N), "tasks": int(snap.T), "groups": int(snap.pod_affinity["n_groups"]), "kinds": int(snap.pod_affinity["n_kinds"]),
           "allocated": int(st.tasks_allocated), "scans": int(st.scans), "kernel_launches": int(st.kernel_launches), "gpu_ms": float(st.gpu_ms),
           "wall_ms": 1e3 * wall, "pairs_logical": int(st.pairs_logical)}
    print("affinity timing:", json.dumps(rec))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(rec, open(os.path.join(out, "affinity_big_session.json"), "w"))


@pytest.mark.parametrize("seed", range(24))
def test_host_level_anti_affinity_runs_on_the_pipeline(eng, seed):
    """Required anti-affinity on kubernetes.io/hostname only (pending and running pods): the host build turns the counter groups into
    atoms of the node's port words (kb_build.h), so the session runs on cycle_kernel when the record geometry allows — bit-exact
    against the object-level oracle, allocate and allocate + backfill."""
    pg = seed % 3 != 2
    snap = aff_gen.host_spread_session(700 + seed, n_nodes=3 + seed % 14, n_groups=3 + seed % 9, pipe_geometry=pg, ports=seed % 4 == 0).flatten(W=2 if pg else 1)
    if snap.pod_affinity is None:
        pytest.skip("no affinity terms drawn")
    for ci, conf in enumerate((PluginConf.default(), PluginConf.from_names([["gang"], ["predicates"]]))):
        for actions in (1, 3):
            o = kbo.allocate(snap, conf, actions=actions)
            eng.load(snap, conf)
            r = eng.allocate()
            assert r.stats.pipeline == (1 if pg else 0)
            if actions & 2:
                r = eng.backfill()
            util.assert_same_decisions(o.decisions, r.decisions, f"seed {seed} conf {ci} actions {actions}")
            util.assert_same_state(o, eng.node_state(), eng.order_state(), f"seed {seed} conf {ci} actions {actions}")


@pytest.mark.parametrize("name,atoms", [("c2", 1), ("c2", 0), ("c3", 1), ("c3", 0)])
def test_one_replica_per_host_at_baseline_size(eng, name, atoms):
    """BASELINE configs 2 / 3 with a tenth of the PodGroups under "one replica per host" (required anti-affinity on
    kubernetes.io/hostname against their own label, synth.add_host_spread).  atoms = 1: the groups become atoms of the port words and
    the cycle runs on cycle_kernel; atoms = 0 (KB_AFF_ATOMS=0): the counter path on the per-visit kernels, where such classes keep
    multi-task runs (ClassAff.pred_multi_ok).  Engine vs the emulation (the object-level oracle is O(pods) per pair)."""
    import json, os
    from kube_batch_b200 import synth
    s, conf = synth.make(name)
    synth.add_host_spread(s, 0.1)
    e = util.emu_allocate(s, conf, mode=1)
    os.environ["KB_AFF_ATOMS"] = str(atoms)
    try:
        eng.load(s, conf)
    finally:
        os.environ.pop("KB_AFF_ATOMS", None)
    eng.allocate()
    r = eng.allocate()
    assert r.stats.pipeline == atoms
    util.assert_same_decisions(e.decisions, r.decisions, f"{name} + host spread vs emulation")
    d = r.decisions
    fb = s.pod_affinity["task_forbid"][:s.T]
    for lab in range(8):
        nodes = d["node"][(fb == np.uint64(3 << (2 * lab))) & (d["kind"] == abi.KB_KIND_ALLOCATED)]
        assert len(nodes) == len(set(nodes.tolist())), "two replicas of one label on a host"
    st = r.stats
    rec = {"workload": name, "path": "cycle_kernel (groups as port-word atoms)" if atoms else "visit_kernel<0,1> (counters)",
           "spread_tasks": int(s.meta["spread_tasks"]), "gpu_ms": float(st.gpu_ms), "kernel_launches": int(st.kernel_launches),
           "allocated": int(st.tasks_allocated), "pairs_logical": int(st.pairs_logical), "pairs_per_s": float(st.pairs_logical) / (st.gpu_ms * 1e-3)}
    print("host spread timing:", json.dumps(rec))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(rec, open(os.path.join(out, f"affinity_host_spread_{name}_{'atoms' if atoms else 'counters'}.json"), "w"))


@pytest.mark.parametrize("seed", range(16))
def test_inter_pod_terms_together_with_preferred_node_affinity(eng, seed):
    """both priorities with a cross-node reduction in one session: InterPodAffinityPriority (min / max of the counts) and
    NodeAffinityPriority (max count) share the passes of aff_prepass_kernel"""
    sb = aff_gen.random_affinity_session(300 + seed, n_nodes=4 + seed % 13, n_groups=3 + seed % 6, node_pref=True, p_affine=0.6 if seed % 3 else 0.0)
    snap = sb.flatten()
    confs = tpa.AFF_CONFS + [PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "-3", "podaffinity.weight": "2"}})]
    for ci, conf in enumerate(confs):
        run_and_check(eng, snap, conf, f"seed {seed} conf {ci}", actions=3)


@pytest.mark.parametrize("seed", range(12))
def test_predicate_step_10_fit_matrix_on_the_gpu(eng, seed):
    """kb_predicate_score (fit only) against the loaded state: the engine's counters + aff_pred vs the oracle's walk over the pods,
    every pending task x every node — predicate step 10 in isolation (predicates.go:1261-1572)."""
    sb = aff_gen.random_affinity_session(800 + seed, n_nodes=6 + seed % 11, n_groups=4 + seed % 5)
    snap = sb.flatten()
    if snap.pod_affinity is None:
        pytest.skip("no affinity terms drawn")
    conf = PluginConf.default()
    eng.load(snap, conf)
    fit, _ = eng.predicate_score(0, snap.T, want_score=False)
    for t in range(snap.T):
        ofit, _ = kbo.predicate_score(snap, conf, t)
        np.testing.assert_array_equal(fit[t], ofit, err_msg=f"seed {seed} task {t}")


@pytest.mark.parametrize("seed", range(12))
def test_the_shipped_action_list_with_host_level_anti_affinity_on_the_gpu(eng, seed):
    """kb_cycle("reclaim, allocate, backfill, preempt") through the C ABI on sessions whose pending pods carry "one replica per host"
    (no placed pod is a member): evict kernels + cycle_kernel / visit_kernel read the member bits from the node records."""
    from test_evict_parity import tier_variants
    s = aff_gen.evict_spread_cluster(40 + seed)
    if s.pod_affinity is None:
        pytest.skip("no spread group drawn")
    acts = ("reclaim", "allocate", "backfill", "preempt")
    for tname, tiers in tier_variants():
        o, ev, order = kbo.cycle(s, tiers, actions=acts, running=s.meta["running"])
        eng.load(s, tiers)
        eng.load_running(s.meta["running"])
        r, gev, gorder, _ = eng.cycle(acts)
        np.testing.assert_array_equal(ev, gev, err_msg=f"seed {seed} {tname}: evicted set")
        np.testing.assert_array_equal(order, gorder, err_msg=f"seed {seed} {tname}: eviction order")
        util.assert_same_decisions(o.decisions, r.decisions, f"seed {seed} {tname}")
        util.assert_same_state(o, eng.node_state(), eng.order_state(), f"seed {seed} {tname}")
    # pods already running are members of the groups: the outcome is withheld iff a member gets evicted (KB_RUNNING_AFF_MEMBER)
    s = aff_gen.evict_spread_cluster(seed, members_running=True)
    if s.pod_affinity is None:
        return
    member = (s.meta["running"]["flags"] & abi.KB_RUNNING_AFF_MEMBER) != 0
    for tname, tiers in tier_variants():
        o, ev, order = kbo.cycle(s, tiers, actions=acts, running=s.meta["running"])
        eng.load(s, tiers)
        eng.load_running(s.meta["running"])
        try:
            r, gev, gorder, _ = eng.cycle(acts)
        except engine.KbError as ex:
            assert ex.code == abi.KB_E_UNSUPPORTED_FEATURE and "withheld" in str(ex)
            continue
        np.testing.assert_array_equal(ev, gev, err_msg=f"members seed {seed} {tname}: evicted set")
        util.assert_same_decisions(o.decisions, r.decisions, f"members seed {seed} {tname}")
        assert not (gev & member).any()
