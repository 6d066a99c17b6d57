"""Seeded random sessions whose pods carry inter-pod (anti)affinity terms (object level: builder.Pod / PodAffinityTerm).
Used by the emulation / GPU parity tests of predicate step 10 and InterPodAffinityPriority."""
from __future__ import annotations

import numpy as np

from kube_batch_b200 import builder as B

HOST, ZONE, RACK = "kubernetes.io/hostname", "zone", "rack"
APPS = ["web", "db", "cache", "batch"]


def _term(rng, hard: bool) -> B.PodAffinityTerm:
    key = [HOST, ZONE, RACK][int(rng.integers(0, 3))] if hard or rng.random() < 0.9 else ""
    t = B.PodAffinityTerm(key)
    r = rng.random()
    if r < 0.5:
        t.match_labels = {"app": APPS[int(rng.integers(0, len(APPS)))]}
    elif r < 0.65:
        t.match_expressions = [("app", "In", [APPS[int(rng.integers(0, 4))], APPS[int(rng.integers(0, 4))]])]
    elif r < 0.75:
        t.match_expressions = [("app", "NotIn", [APPS[int(rng.integers(0, 4))]])]
    elif r < 0.83:
        t.match_expressions = [("tier", "Exists", [])]
    elif r < 0.9:
        t.match_expressions = [("tier", "DoesNotExist", [])]
        t.match_labels = {"app": APPS[int(rng.integers(0, 4))]}
    elif r < 0.95:
        pass                                    # empty selector: everything in the namespaces
    else:
        t.nil_selector = True                   # nothing
    r = rng.random()
    if r < 0.15:
        t.namespaces = ["ns1", "ns2"]
    elif r < 0.25:
        t.namespaces = ["ns2"]
    return t


def _spec(rng, p_req=0.6, p_pref=0.5):
    a = B.PodAffinity()
    if rng.random() < p_req:
        a.required = [_term(rng, True) for _ in range(1 if rng.random() < 0.8 else 2)]
    if rng.random() < p_pref:
        a.preferred = [(int(rng.integers(1, 100)), _term(rng, False)) for _ in range(int(rng.integers(1, 3)))]
    return a


def random_affinity_session(seed: int, n_nodes: int = 12, n_groups: int = 6, p_affine: float = 0.6, besteffort: bool = False,
                            spec_pool: int = 0) -> B.SessionBuilder:
    """spec_pool > 0: the PodGroups draw their (labels, affinity, anti-affinity) from that many templates (large sessions stay
    within the 64 counter groups of kb_pod_affinity)."""
    rng = np.random.default_rng(seed)
    pool = []
    for _ in range(spec_pool):
        lab = {"app": APPS[int(rng.integers(0, 4))]}
        pool.append((lab, _spec(rng) if rng.random() < p_affine * 0.6 else None, _spec(rng) if rng.random() < p_affine else None))
    sb = B.SessionBuilder()
    nq = int(rng.integers(1, 3))
    for q in range(nq):
        sb.add_queue(B.Queue(f"q{q}", weight=int(rng.integers(1, 4)), creation=q))
    zones = int(rng.integers(2, 4))
    for i in range(n_nodes):
        labels = {HOST: f"n{i:03d}"}
        if rng.random() < 0.9:
            labels[ZONE] = f"z{i % zones}"
        if rng.random() < 0.7:
            labels[RACK] = f"r{i % 4}"
        cpu = int(rng.choice([4, 8, 16]))
        sb.add_node(B.build_node(f"n{i:03d}", B.build_resource_list(str(cpu), f"{cpu * 2}Gi"), labels=labels, pods=int(rng.choice([4, 110]))))
    uid = 0
    # pods already running: some belong to session jobs (util.PodLister sees them), some to a group without queue (NodeInfo.Tasks only)
    sb.add_pod_group(B.PodGroup("ns1", "run-a", "q0", min_member=1))
    sb.add_pod_group(B.PodGroup("ns2", "run-b", "nosuchqueue", min_member=1))
    load = {}
    for i in range(int(rng.integers(0, n_nodes))):
        ns, grp = ("ns1", "run-a") if rng.random() < 0.7 else ("ns2", "run-b")
        host = int(rng.integers(0, n_nodes))
        if load.get(host, 0) >= 2:
            continue                      # never over-commit a node: the cache would refuse the pod (node_info.go:161-167)
        load[host] = load.get(host, 0) + 1
        p = B.build_pod(ns, f"r{i}", f"n{host:03d}", "Running", B.build_resource_list("1", "1Gi"), grp,
                        labels={"app": APPS[int(rng.integers(0, 4))]})
        if rng.random() < 0.3:
            p.labels["tier"] = "x"
        if rng.random() < 0.5:
            if rng.random() < 0.5:
                p.pod_affinity = _spec(rng, 0.5, 0.6)
            if rng.random() < 0.6:
                p.pod_anti_affinity = _spec(rng, 0.5, 0.6)
        p.uid = f"u{uid:05d}"; uid += 1
        sb.add_pod(p)
    for g in range(n_groups):
        ns = "ns1" if rng.random() < 0.7 else "ns2"
        ntask = int(rng.integers(1, 6))
        sb.add_pod_group(B.PodGroup(ns, f"pg{g}", f"q{int(rng.integers(0, nq))}", min_member=int(rng.integers(1, ntask + 1)), creation=g))
        labels = {"app": APPS[int(rng.integers(0, 4))]}
        if rng.random() < 0.3:
            labels["tier"] = "x"
        aff = _spec(rng) if rng.random() < p_affine * 0.6 else None
        anti = _spec(rng) if rng.random() < p_affine else None
        if pool:
            labels, aff, anti = pool[int(rng.integers(0, len(pool)))]
        req = {} if (besteffort and rng.random() < 0.5) else B.build_resource_list(str(int(rng.choice([1, 2, 3]))), f"{int(rng.choice([1, 2, 4]))}Gi")
        for i in range(ntask):
            p = B.build_pod(ns, f"pg{g}-{i}", "", "Pending", req, f"pg{g}", labels=dict(labels))
            p.pod_affinity, p.pod_anti_affinity = aff, anti
            p.creation = int(rng.integers(0, 4))
            p.uid = f"u{uid:05d}"; uid += 1
            if besteffort and not req and rng.random() < 0.3:
                p.requests = {}
            sb.add_pod(p)
    return sb
