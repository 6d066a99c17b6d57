"""Inter-pod (anti)affinity: predicate step 10 (vendor/.../predicates/predicates.go:1261-1572, slow path) and
InterPodAffinityPriority (vendor/.../priorities/interpod_affinity.go:99-235).

The reference's unit tests do not cover these vendored functions (their _test.go files are stripped): the vectors below are
hand-computed from the Go source and pin the oracle; the emulation of the device algorithm (aggregated counters per topology
domain, kb_aff.h) is then checked against the oracle, which walks the raw pod objects."""
import numpy as np
import pytest

from kube_batch_b200 import abi, builder as B
from kube_batch_b200.snapshot import PluginConf
from oracle import kbo
import aff_gen
import util

HOST, ZONE = "kubernetes.io/hostname", "zone"
ONLY_PODAFF = PluginConf.from_names([["gang"], ["predicates", "nodeorder"]],
                                    {"nodeorder": {"leastrequested.weight": "0", "balancedresource.weight": "0", "nodeaffinity.weight": "0"}})
AFF_CONFS = [PluginConf.default(),
             PluginConf.from_names([["priority", "gang"], ["drf", "predicates", "proportion", "nodeorder"]],
                                   {"nodeorder": {"podaffinity.weight": "-3", "leastrequested.weight": "0"}}),
             PluginConf.from_names([["gang"], ["predicates"]]),
             PluginConf.from_names([["gang", "drf"], ["nodeorder"]], {"nodeorder": {"podaffinity.weight": "5"}})]


def cluster(n_nodes=4, zones=2, cpu="8"):
    sb = B.SessionBuilder()
    sb.add_queue(B.Queue("q1"))
    for i in range(n_nodes):
        sb.add_node(B.build_node(f"n{i}", B.build_resource_list(cpu, "16Gi"), labels={HOST: f"n{i}", ZONE: f"z{i % zones}"}, pods=110))
    sb.add_pod_group(B.PodGroup("ns", "pg1", "q1", min_member=1))
    sb.add_pod_group(B.PodGroup("ns", "run", "q1", min_member=1))
    return sb


def pod(name, labels, node="", phase="Pending", group="pg1", ns="ns", creation=0, cpu="1"):
    p = B.build_pod(ns, name, node, phase, B.build_resource_list(cpu, "1Gi"), group, labels=labels)
    p.creation = creation
    return p


def term(key, **labels):
    return B.PodAffinityTerm(key, match_labels=dict(labels))


def both(snap, conf, actions=1):
    o = kbo.allocate(snap, conf, actions=actions)
    e = util.emu_allocate(snap, conf, actions=actions, mode=1)
    util.assert_same_decisions(o.decisions, e.decisions, "emulation vs oracle")
    return o


# ---------------- predicate: hand-computed ----------------
def test_self_anti_affinity_on_hostname_places_one_pod_per_node():
    sb = cluster(4)
    for i in range(5):
        p = pod(f"p{i}", {"app": "web"}, creation=i)
        p.pod_anti_affinity = B.PodAffinity(required=[term(HOST, app="web")])
        sb.add_pod(p)
    o = both(sb.flatten(), PluginConf.default())
    assert sorted(o.decisions["node"][:4].tolist()) == [0, 1, 2, 3]
    assert o.decisions["kind"][4] == abi.KB_KIND_NONE          # every host holds a web pod


def test_required_affinity_follows_the_zone_of_an_existing_pod():
    sb = cluster(4, zones=2)
    sb.add_pod(pod("db0", {"app": "db"}, node="n1", phase="Running", group="run"))       # n1 is in zone z1 (with n3)
    for i in range(3):
        p = pod(f"p{i}", {"app": "web"}, creation=i)
        p.pod_affinity = B.PodAffinity(required=[term(ZONE, app="db")])
        sb.add_pod(p)
    o = both(sb.flatten(), PluginConf.default())
    assert set(o.decisions["node"].tolist()) <= {1, 3}
    assert (o.decisions["kind"] == abi.KB_KIND_ALLOCATED).all()


def test_first_pod_of_a_self_affine_series_passes_then_the_zone_is_fixed():
    """predicates.go:1545-1560: no pod anywhere matches the terms and the pod matches its own terms -> the check passes."""
    sb = cluster(6, zones=3)
    for i in range(4):
        p = pod(f"p{i}", {"app": "ring"}, creation=i)
        p.pod_affinity = B.PodAffinity(required=[term(ZONE, app="ring")])
        sb.add_pod(p)
    o = both(sb.flatten(), PluginConf.default())
    first = int(o.decisions["node"][0])
    assert first >= 0 and all(int(n) % 3 == first % 3 for n in o.decisions["node"])     # all in the first pod's zone
    # a pod that does NOT match its own terms never gets the escape
    sb = cluster(6, zones=3)
    p = pod("lonely", {"app": "other"})
    p.pod_affinity = B.PodAffinity(required=[term(ZONE, app="ring")])
    sb.add_pod(p)
    o = both(sb.flatten(), PluginConf.default())
    assert o.decisions["kind"][0] == abi.KB_KIND_NONE


def test_anti_affinity_of_an_existing_pod_rejects_the_incoming_pod():
    """satisfiesExistingPodsAntiAffinity (:1400-1439): the INCOMING pod has no terms at all."""
    sb = cluster(4, zones=2)
    guard = pod("guard", {"app": "db"}, node="n0", phase="Running", group="run")
    guard.pod_anti_affinity = B.PodAffinity(required=[term(ZONE, app="web")])
    sb.add_pod(guard)
    for i in range(3):
        sb.add_pod(pod(f"w{i}", {"app": "web"}, creation=i))
    sb.add_pod(pod("other", {"app": "cache"}, creation=9))
    snap = sb.flatten()
    assert snap.flags & abi.KB_SNAPSHOT_PLACED_POD_AFFINITY
    o = both(snap, PluginConf.default())
    assert set(o.decisions["node"][:3].tolist()) <= {1, 3}       # zone z0 (n0, n2) is closed for web pods
    fit, _ = kbo.predicate_score(snap, PluginConf.default(), 3)
    assert fit.tolist() == [1, 1, 1, 1]                         # ... but not for the cache pod


def test_a_term_without_namespaces_only_sees_the_owners_namespace():
    sb = cluster(2, zones=2)
    sb.add_pod_group(B.PodGroup("other", "pgo", "q1", min_member=1))
    sb.add_pod(pod("web-elsewhere", {"app": "web"}, node="n0", phase="Running", group="run", ns="ns"))
    p = pod("p0", {"app": "x"}, ns="other", group="pgo")
    p.pod_anti_affinity = B.PodAffinity(required=[term(HOST, app="web")])       # namespaces empty -> {"other"}
    sb.add_pod(p)
    q = pod("p1", {"app": "x"}, ns="other", group="pgo", creation=1)
    t = term(HOST, app="web"); t.namespaces = ["ns"]
    q.pod_anti_affinity = B.PodAffinity(required=[t])
    sb.add_pod(q)
    snap = sb.flatten()
    conf = PluginConf.default()
    assert kbo.predicate_score(snap, conf, 0)[0].tolist() == [1, 1]
    assert kbo.predicate_score(snap, conf, 1)[0].tolist() == [0, 1]
    both(snap, conf)


def test_slow_path_anti_affinity_needs_a_pod_that_matches_all_terms():
    """podMatchesPodAffinityTerms (:1296-1320) is called with the whole term list: an existing pod rejects the node only if it
    matches the namespaces + selector of EVERY anti-affinity term (and shares every topology) — the meta == nil path kube-batch runs."""
    sb = cluster(2, zones=1)
    sb.add_pod(pod("a", {"app": "a"}, node="n0", phase="Running", group="run"))
    sb.add_pod(pod("ab", {"app": "a", "tier": "b"}, node="n1", phase="Running", group="run"))
    p = pod("p0", {"app": "x"})
    p.pod_anti_affinity = B.PodAffinity(required=[term(HOST, app="a"), term(HOST, tier="b")])
    sb.add_pod(p)
    snap = sb.flatten()
    assert kbo.predicate_score(snap, PluginConf.default(), 0)[0].tolist() == [1, 0]
    both(snap, PluginConf.default())


def test_node_without_the_topology_label_never_matches():
    sb = cluster(3, zones=1)
    sb.nodes[2].labels.pop(ZONE)
    sb.add_pod(pod("db", {"app": "db"}, node="n2", phase="Running", group="run"))
    p = pod("p0", {"app": "web"})
    p.pod_affinity = B.PodAffinity(required=[term(ZONE, app="db")])
    sb.add_pod(p)
    snap = sb.flatten()
    # the only db pod sits on a node without a zone label: NodesHaveSameTopologyKey is false everywhere, and the pod exists,
    # so the first-of-series escape does not apply either
    assert kbo.predicate_score(snap, PluginConf.default(), 0)[0].tolist() == [0, 0, 0]
    both(snap, PluginConf.default())


# ---------------- priority: hand-computed ----------------
def test_preferred_affinity_scores_hand_computed():
    sb = cluster(3, zones=2)                       # n0, n2 in z0; n1 in z1
    sb.add_pod(pod("db", {"app": "db"}, node="n0", phase="Running", group="run"))
    p = pod("p0", {"app": "web"})
    p.pod_affinity = B.PodAffinity(preferred=[(10, term(ZONE, app="db"))])
    sb.add_pod(p)
    snap = sb.flatten()
    fit, score = kbo.predicate_score(snap, ONLY_PODAFF, 0)
    # counts: n0 10, n1 0, n2 10 -> min 0 max 10 -> 10, 0, 10
    assert fit.tolist() == [1, 1, 1] and score.tolist() == [10.0, 0.0, 10.0]
    both(snap, ONLY_PODAFF)


def test_negative_counts_and_truncation_hand_computed():
    sb = cluster(4, zones=4)
    # existing pods: a "noisy" pod on n0 that the incoming pod avoids (weight 4), an admirer on n1 whose preferred affinity
    # selects the incoming pod (weight 1), two admirers on n2 with a REQUIRED affinity term (hard weight 1 each) + one preferred 1
    sb.add_pod(pod("noisy", {"app": "noisy"}, node="n0", phase="Running", group="run"))
    adm = pod("adm1", {"app": "fan"}, node="n1", phase="Running", group="run")
    adm.pod_affinity = B.PodAffinity(preferred=[(1, term(HOST, app="web"))])
    sb.add_pod(adm)
    for i in range(2):
        a = pod(f"adm2-{i}", {"app": "fan"}, node="n2", phase="Running", group="run")
        a.pod_affinity = B.PodAffinity(required=[term(HOST, app="web")], preferred=[(1, term(HOST, app="web"))] if i == 0 else [])
        sb.add_pod(a)
    p = pod("p0", {"app": "web"})
    p.pod_anti_affinity = B.PodAffinity(preferred=[(4, term(HOST, app="noisy"))])
    sb.add_pod(p)
    snap = sb.flatten()
    fit, score = kbo.predicate_score(snap, ONLY_PODAFF, 0)
    # required affinity of the admirers on n2 towards app=web: no web pod exists, they do not constrain p0 (only THEIR placement)
    # counts: n0 -4, n1 +1, n2 +1+1+1 = 3, n3 0 -> min -4, max 3 -> 10*(c+4)/7 = 0, 7.14 -> 7, 10, 5.71 -> 5
    assert fit.tolist() == [1, 1, 1, 1] and score.tolist() == [0.0, 7.0, 10.0, 5.0]
    both(snap, ONLY_PODAFF)


def test_only_pods_on_feasible_nodes_count():
    """nodeNameToInfo holds the FEASIBLE nodes only (util/scheduler_helper.go:219-230): a matching pod on a node the task cannot use
    adds nothing, not even to the other nodes of its zone."""
    sb = cluster(3, zones=1, cpu="2")
    sb.add_pod(pod("db", {"app": "db"}, node="n0", phase="Running", group="run", cpu="2"))     # n0 is full
    p = pod("p0", {"app": "web"})
    p.pod_affinity = B.PodAffinity(preferred=[(5, term(ZONE, app="db"))])
    sb.add_pod(p)
    snap = sb.flatten()
    fit, score = kbo.predicate_score(snap, ONLY_PODAFF, 0)
    assert fit.tolist() == [0, 1, 1] and score.tolist() == [0.0, 0.0, 0.0]
    both(snap, ONLY_PODAFF)


def test_a_pod_placed_this_session_is_located_through_the_first_unbound_pod():
    """cachedNodeInfo.GetNodeInfo (plugins/nodeorder/nodeorder.go:49-63): the pod object of a task placed in this session still has
    an empty Spec.NodeName, so its node is looked up as "the first node holding ANY pod with an empty node name" (ascending node
    order is the deterministic rule).  c0 lands on n0, a0 on n2; b0 prefers a0's HOST — and the weight goes to n0."""
    sb = cluster(3, zones=3)
    sb.add_pod_group(B.PodGroup("ns", "pg0", "q1", min_member=1, creation=0))
    sb.add_pod_group(B.PodGroup("ns", "pg2", "q1", min_member=1, creation=2))
    sb.pod_groups[0].creation = 1                                        # pg1 second
    c0 = pod("c0", {"app": "c"}, group="pg0")
    c0.node_selector = {HOST: "n0"}
    a0 = pod("a0", {"app": "a"}, group="pg1")
    a0.node_selector = {HOST: "n2"}
    b0 = pod("b0", {"app": "b"}, group="pg2")
    b0.pod_affinity = B.PodAffinity(preferred=[(7, term(HOST, app="a"))])
    for p in (c0, a0, b0):
        sb.add_pod(p)
    snap = sb.flatten()
    o = both(snap, ONLY_PODAFF)
    names = snap.meta["tasks"]
    where = {names[t]: int(o.decisions["node"][t]) for t in range(snap.T)}
    assert where["ns/c0"] == 0 and where["ns/a0"] == 2
    assert where["ns/b0"] == 0, "the weight of a0 (really on n2) is credited to n0, the first node with a not-yet-bound pod"


# ---------------- emulation of the device algorithm vs the oracle ----------------
@pytest.mark.parametrize("seed", range(48))
def test_random_affinity_sessions_emulation_matches_the_oracle(seed):
    sb = aff_gen.random_affinity_session(seed, n_nodes=4 + seed % 13, n_groups=3 + seed % 6, besteffort=seed % 4 == 3)
    snap = sb.flatten()
    if snap.pod_affinity is None:
        pytest.skip("no affinity terms drawn")
    for ci, conf in enumerate(AFF_CONFS):
        o = kbo.allocate(snap, conf, actions=3)
        e = util.emu_allocate(snap, conf, actions=3, mode=1)
        util.assert_same_decisions(o.decisions, e.decisions, f"seed {seed} conf {ci}")
        st = util.emu_states(e)
        util.assert_same_state(o, st[0], st[1], f"seed {seed} conf {ci}")


def test_affinity_terms_without_the_flattened_tables_are_refused():
    sb = cluster(2)
    p = pod("p0", {"app": "web"})
    p.pod_anti_affinity = B.PodAffinity(required=[term(HOST, app="web")])
    sb.add_pod(p)
    snap = sb.flatten()
    assert snap.task_flags[0] & abi.KB_TASK_HAS_POD_AFFINITY
    snap.pod_affinity = None
    with pytest.raises(RuntimeError, match="no kb_pod_affinity"):
        util.emu_allocate(snap, PluginConf.default(), mode=1)


def test_a_listed_pod_on_a_node_outside_the_session_fails_every_predicate():
    """cache.Snapshot drops NotReady nodes (cache/cache.go:633-640) but the jobs keep their tasks: util.PodLister hands the pod to
    InterPodAffinityMatches, CachedNodeInfo.GetNodeInfo does not find its node (plugins/util/util.go:93-100) and the error fails the
    predicate for every (pod, node) pair (vendor/.../predicates.go:1381-1393) — no affinity term anywhere.  KB_SNAPSHOT_LISTED_POD_WITHOUT_NODE."""
    sb = cluster(3)
    sb.add_pod(pod("lost", {"app": "db"}, node="gone-node", phase="Running", group="run"))
    for i in range(3):
        sb.add_pod(pod(f"p{i}", {"app": "web"}, creation=i))
    snap = sb.flatten()
    assert snap.flags & abi.KB_SNAPSHOT_LISTED_POD_WITHOUT_NODE and snap.pod_affinity is None
    o = both(snap, PluginConf.default(), actions=3)
    assert (o.decisions["kind"] == abi.KB_KIND_NONE).all()
    for mode in (0, 5):
        e = util.emu_allocate(snap, PluginConf.default(), actions=3, mode=mode)
        util.assert_same_decisions(o.decisions, e.decisions, f"mode {mode}")
    # without the predicates plugin nobody calls InterPodAffinityMatches
    o = both(snap, PluginConf.from_names([["gang", "priority"], ["drf", "nodeorder"]]), actions=3)
    assert (o.decisions["kind"] == abi.KB_KIND_ALLOCATED).all()


def test_synthetic_host_spread_tables_are_what_the_flattener_produces():
    """synth.add_host_spread writes kb_pod_affinity directly (BASELINE-size sessions are not built from objects): same structure as
    builder.flatten_pod_affinity gives for pods that carry {app=L} + required anti-affinity {app=L, hostname}."""
    from kube_batch_b200 import synth
    sb = cluster(4)
    for i in range(3):
        p = pod(f"p{i}", {"app": "web"}, creation=i)
        p.pod_anti_affinity = B.PodAffinity(required=[term(HOST, app="web")])
        sb.add_pod(p)
    a = sb.flatten().pod_affinity
    s, _ = synth.make("c1")
    b = synth.add_host_spread(s, frac=1.0, labels=1).pod_affinity
    assert (a["n_groups"], a["n_keysets"], a["n_kinds"]) == (b["n_groups"], b["n_keysets"], b["n_kinds"]) == (2, 1, 0)
    assert a["task_forbid"].tolist() == [3, 3, 3] and set(b["task_forbid"][: s.T].tolist()) == {3}
    assert a["task_contrib"].tolist() == [3, 3, 3] and np.array_equal(b["task_contrib"], b["task_forbid"])
    assert a["node_domain"].tolist() == [[0, 1, 2, 3]] and b["node_domain"].tolist() == [list(range(s.N))]
    e = util.emu_allocate(s, PluginConf.default(), mode=1)
    d = e.decisions
    placed = d["node"][d["kind"] == abi.KB_KIND_ALLOCATED]
    assert len(placed) == len(set(placed.tolist())) > 0


@pytest.mark.parametrize("seed", range(40))
def test_host_level_anti_affinity_as_atoms_in_every_launch_mode(seed):
    """Sessions whose only inter-pod constraint is required anti-affinity on kubernetes.io/hostname: kb_build.h encodes the counter
    groups as atoms of the node's port words (ClassRec.port_conflict / aff_own), so they run in EVERY launch mode — overlap (0),
    plain (1), the persistent pipeline's stale-list + patch protocol (5) — and must still equal the object-level oracle.
    KB_AFF_ATOMS=0 keeps them on the counter path; both must agree."""
    import os
    pg = seed % 2 == 0
    snap = aff_gen.host_spread_session(seed, n_nodes=3 + seed % 14, n_groups=3 + seed % 9, pipe_geometry=pg, ports=seed % 3 == 0).flatten(W=2 if pg else 1)
    if snap.pod_affinity is None:
        pytest.skip("no affinity terms drawn")
    for ci, conf in enumerate((PluginConf.default(), PluginConf.from_names([["gang"], ["predicates"]]),
                               PluginConf.from_names([["priority", "gang"], ["drf", "predicates", "proportion", "nodeorder"]], {"nodeorder": {"podaffinity.weight": "0"}}))):
        o = kbo.allocate(snap, conf, actions=3)
        for mode in (0, 1, 5):
            e = util.emu_allocate(snap, conf, actions=3, mode=mode)
            util.assert_same_decisions(o.decisions, e.decisions, f"seed {seed} conf {ci} mode {mode}")
            st = util.emu_states(e)
            util.assert_same_state(o, st[0], st[1], f"seed {seed} conf {ci} mode {mode}")
        os.environ["KB_AFF_ATOMS"] = "0"
        try:
            e = util.emu_allocate(snap, conf, actions=3, mode=1)
        finally:
            os.environ.pop("KB_AFF_ATOMS", None)
        util.assert_same_decisions(o.decisions, e.decisions, f"seed {seed} conf {ci} counter path")


@pytest.mark.parametrize("seed", range(32))
def test_inter_pod_terms_together_with_preferred_node_affinity(seed):
    sb = aff_gen.random_affinity_session(300 + seed, n_nodes=4 + seed % 13, n_groups=3 + seed % 6, node_pref=True, p_affine=0.6 if seed % 3 else 0.0)
    snap = sb.flatten()
    confs = AFF_CONFS + [PluginConf.from_names([["gang"], ["predicates", "nodeorder"]], {"nodeorder": {"nodeaffinity.weight": "-3", "podaffinity.weight": "2"}})]
    for ci, conf in enumerate(confs):
        o = kbo.allocate(snap, conf, actions=3)
        e = util.emu_allocate(snap, conf, actions=3, mode=1)
        util.assert_same_decisions(o.decisions, e.decisions, f"seed {seed} conf {ci}")


@pytest.mark.parametrize("seed", range(24))
def test_reclaim_and_preempt_with_host_level_anti_affinity_on_pending_pods(seed):
    """The shipped action list on sessions whose PENDING pods carry "one replica per host" while no placed pod is a member of a counter
    group: a victim (a Running pod) is then never a member, an eviction changes no member bit, the preemptors are Pipelined (not
    listed by util.PodLister) — the atoms stay exact through reclaim / allocate / backfill / preempt.  Emulation vs the oracle (whose
    PodLister follows every status change)."""
    from test_evict_parity import compare, tier_variants
    s = aff_gen.evict_spread_cluster(seed)
    if s.pod_affinity is None:
        pytest.skip("no spread group drawn")
    for tname, tiers in tier_variants():
        for acts in (("reclaim", "allocate", "backfill", "preempt"), ("reclaim",), ("allocate", "preempt")):
            o, ev, order = kbo.cycle(s, tiers, actions=acts, running=s.meta["running"])
            g, gev, gorder = util.emu_cycle(s, tiers, acts, s.meta["running"], mode=1)
            compare(f"seed {seed} {tname} {acts}", o, ev, order, g, gev, gorder, util.emu_states(g))


@pytest.mark.parametrize("seed", range(16))
def test_evicting_a_member_of_a_counter_group_withholds_the_outcome(seed):
    """The pods already running carry the labels + terms too, so a victim can be a MEMBER of a counter group: its eviction takes it out of
    util.PodLister and would have to clear a member bit of its node record.  The engine does not track that; it notices the eviction
    (KB_RUNNING_AFF_MEMBER, also inside a Statement that is discarded later) and withholds the cycle's outcome
    (KB_E_UNSUPPORTED_FEATURE -> the shim reruns the cycle with the original actions).  Otherwise the outcome equals the oracle's and
    no member is among the evicted."""
    from test_evict_parity import compare, tier_variants
    s = aff_gen.evict_spread_cluster(seed, members_running=True)
    if s.pod_affinity is None:
        pytest.skip("no spread group drawn")
    member = (s.meta["running"]["flags"] & abi.KB_RUNNING_AFF_MEMBER) != 0
    for tname, tiers in tier_variants():
        for acts in (("reclaim", "allocate", "backfill", "preempt"), ("preempt",)):
            o, ev, order = kbo.cycle(s, tiers, actions=acts, running=s.meta["running"])
            try:
                g, gev, gorder = util.emu_cycle(s, tiers, acts, s.meta["running"], mode=1)
            except RuntimeError as ex:
                assert "withheld" in str(ex)
                continue
            compare(f"seed {seed} {tname} {acts}", o, ev, order, g, gev, gorder, util.emu_states(g))
            assert not (ev & member).any()
