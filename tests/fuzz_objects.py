"""Object-level fuzz: random clusters built with kube_batch_b200.builder (nodes with labels / taints / pressure flags / pod caps,
Running / terminating / Succeeded / Failed pods, queues, PodGroups with minMember and priorities, pending pods with selectors,
required node affinity, tolerations, host ports, init containers, best-effort and sub-epsilon requests, extra scalar resources)
-> CPU emulation of the device algorithm in every launch mode (overlap protocol, plain, chained K=2/4) x {allocate, allocate+backfill}
vs the oracle.  Found in round 1: over-committed nodes must not reach the engine (the reference cache refuses them) and the
phantom-Allocated corner of backfill.  usage: python tests/fuzz_objects.py [first_seed [count]]"""
import os, sys, time
import numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
from kube_batch_b200 import builder as B, abi
from kube_batch_b200.snapshot import PluginConf, PluginOption
from oracle import kbo
import util
from test_emu_parity import CONFS

def rand_session(seed):
    rng = np.random.default_rng(seed)
    b = B.SessionBuilder()
    nq = int(rng.integers(1, 5))
    for q in range(nq):
        b.add_queue(B.Queue(f"q{q}", int(rng.integers(1, 5)), creation=int(rng.integers(0, 3))))
    nn = int(rng.integers(1, 40))
    zones = ["a", "b", "c"] if rng.random() < 0.7 else [f"z{i}" for i in range(int(rng.integers(4, 90)))]
    gpu = rng.random() < 0.5
    for n in range(nn):
        alloc = {"cpu": float(rng.choice([2, 4, 8, 16])), "memory": float(rng.choice([4, 8, 16, 64])) * 1e9, "pods": int(rng.choice([0, 2, 5, 110]))}
        if gpu and rng.random() < 0.5:
            alloc["nvidia.com/gpu"] = float(rng.choice([1, 2, 8]))
        if rng.random() < 0.3:
            alloc["example.com/foo"] = float(rng.choice([1, 4]))
        taints = []
        if rng.random() < 0.2: taints.append(("dedicated", str(rng.choice(["x", "y"])), str(rng.choice(["NoSchedule", "NoExecute", "PreferNoSchedule"]))))
        b.add_node(B.Node(f"n{n:03d}", alloc, labels={"zone": str(rng.choice(zones)), "rank": str(n % 7)}, taints=taints,
                          unschedulable=bool(rng.random() < 0.05), ready=None if rng.random() < 0.9 else False,
                          memory_pressure=bool(rng.random() < 0.05), disk_pressure=bool(rng.random() < 0.05)))
    ng = int(rng.integers(1, 15))
    pid = 0
    for g in range(ng):
        b.add_pod_group(B.PodGroup("ns", f"g{g:02d}", f"q{int(rng.integers(0, nq))}", min_member=int(rng.integers(0, 5)),
                                   priority=int(rng.integers(0, 3)), creation=int(rng.integers(0, 4))))
        hetero = rng.random() < 0.3
        base = {"cpu": float(rng.choice([0.25, 0.5, 1, 2, 4])), "memory": float(rng.choice([0.5, 1, 2, 8])) * 1e9}
        if gpu and rng.random() < 0.3: base["nvidia.com/gpu"] = 1.0
        sel = {"zone": str(rng.choice(zones))} if rng.random() < 0.3 else {}
        for k in range(int(rng.integers(1, 9))):
            req = dict(base)
            if hetero: req["cpu"] = float(rng.choice([0.25, 0.5, 1, 2]))
            kind = rng.random()
            kw = dict(group=f"g{g:02d}", creation=int(rng.integers(0, 5)), priority=int(rng.integers(0, 3)) if rng.random() < 0.5 else None)
            if kind < 0.08: req = {}
            elif kind < 0.12: req = {"cpu": 0.005}
            elif kind < 0.16: req = {"example.com/foo": float(rng.choice([1, 2]))}
            if rng.random() < 0.1: kw["init_requests"] = [{"cpu": 3.0}]
            if rng.random() < 0.15: kw["host_ports"] = [(str(rng.choice(["", "10.0.0.1"])), str(rng.choice(["TCP", "UDP"])), int(rng.choice([80, 8080])))]
            if rng.random() < 0.2: kw["tolerations"] = [("dedicated", str(rng.choice(["Equal", "Exists"])), str(rng.choice(["x", "y"])), str(rng.choice(["", "NoSchedule"])))]
            if rng.random() < 0.15: kw["affinity_terms"] = [[("zone", str(rng.choice(["In", "NotIn"])), [str(rng.choice(zones))])], [("rank", str(rng.choice(["Gt", "Lt", "Exists"])), ["3"])]]
            phase = "Pending"; node = ""
            if rng.random() < 0.25:
                phase = str(rng.choice(["Running", "Running", "Running", "Succeeded", "Failed"])); node = f"n{int(rng.integers(0, nn)):03d}"
            b.add_pod(B.Pod("ns", f"p{pid:04d}", node, phase, req, node_selector=dict(sel), deleting=bool(phase == "Running" and rng.random() < 0.3), **kw))
            pid += 1
    return b.flatten()

t0 = time.time(); n = 0; bad = 0; phantoms = 0
confs = list(CONFS.items())
start = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for seed in range(start, start + count):
    try:
        s = rand_session(seed)
    except AssertionError as e:
        continue
    for ci in (seed % len(confs), (seed * 7 + 3) % len(confs)):
        cname, conf = confs[ci]
        for actions in (1, 3):
            try:
                o = kbo.allocate(s, conf, actions=actions)
            except RuntimeError as e:
                print("oracle error", seed, cname, e); continue
            for mode in (0, 1, 2, 4, 5):
                n += 1
                try:
                    e = util.emu_allocate(s, conf, actions=actions, mode=mode)
                    ph = (o.decisions["kind"] == 1) & (o.decisions["node"] == -1)
                    if ph.any():
                        phantoms += 1                 # the phantom-Allocated corner of backfill is compared like everything else now
                    util.assert_same_decisions(o.decisions, e.decisions, f"seed {seed} {cname} a{actions} m{mode}")
                    ns, os_ = util.emu_states(e)
                    util.assert_same_state(o, ns, os_, f"seed {seed} {cname} a{actions} m{mode}")
                except Exception as ex:
                    bad += 1
                    print("MISMATCH", seed, cname, actions, mode, str(ex)[:300])
print(f"{n} comparisons, {bad} bad, {phantoms} with phantom-Allocated tasks (compared), {time.time()-t0:.0f}s")
