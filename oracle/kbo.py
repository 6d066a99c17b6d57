"""ctypes binding of the CPU oracle (oracle/libkboracle.so).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
the product package (kube_batch_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

from kube_batch_b200 import abi
from kube_batch_b200.snapshot import PluginConf, Snapshot

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkboracle.so")

KBO_MODE_OPTIMISED = 0
KBO_MODE_FAITHFUL = 1


class kbo_opts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("threads", C.c_int32), ("max_tasks", C.c_int64), ("max_seconds", C.c_double),
                ("actions", C.c_int32), ("warm_tasks", C.c_int32)]


class kbo_result(C.Structure):
    _fields_ = [
        ("pairs_logical", C.c_uint64), ("tasks_processed", C.c_uint32), ("tasks_allocated", C.c_uint32),
        ("tasks_pipelined", C.c_uint32), ("jobs_ready", C.c_uint32), ("visits", C.c_uint32),
        ("truncated", C.c_uint32), ("evictions", C.c_uint32), ("timed_tasks", C.c_uint32), ("seconds", C.c_double),
    ]


def _stale() -> bool:
    src = os.path.join(_HERE, "kb_oracle.cpp")
    return not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "kb_oracle.h")),
        os.path.getmtime(os.path.join(_HERE, "..", "include", "kbgpu.h")))


def build(force: bool = False) -> str:
    """(Re)build the oracle library if needed.  Several ranks / test workers may call this at once: serialise on a lock
    file and re-check, so that nobody dlopens a half-written .so."""
    if force or _stale():
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            if force or _stale():
                subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.kbo_last_error.restype = C.c_char_p
        L.kbo_share.restype = C.c_double
        L.kbo_share.argtypes = [C.c_double, C.c_double]
        for f in ("kbo_least_requested", "kbo_most_requested", "kbo_balanced"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_int64] * 4
        _lib = L
    return _lib


@dataclass
class OracleOut:
    decisions: np.ndarray      # structured, abi.DECISION_DTYPE
    result: kbo_result
    node_idle: np.ndarray
    node_releasing: np.ndarray
    node_used: np.ndarray
    node_pods: np.ndarray
    node_nz_cpu: np.ndarray
    node_nz_mem: np.ndarray
    node_ports: np.ndarray
    job_share: np.ndarray
    job_ready: np.ndarray
    queue_share: np.ndarray
    queue_deserved: np.ndarray
    queue_allocated: np.ndarray

    def bind_map(self):
        """FakeBinder.Binds (util/test_utils.go:95-112): task -> node for dispatched tasks."""
        d = self.decisions
        return {int(t): int(d["node"][t]) for t in np.nonzero(d["dispatched"])[0]}


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


KBO_ACTION_ALLOCATE = 1
KBO_ACTION_BACKFILL = 2

_PO_FIELDS = [("pod_ns", C.c_int32), ("lab_off", C.c_uint32), ("lab_key", C.c_int32), ("lab_val", C.c_int32), ("has_aff", C.c_uint8),
              ("has_anti", C.c_uint8), ("term_off", C.c_uint32), ("term_kind", C.c_int32), ("term_weight", C.c_int32),
              ("term_topo", C.c_int32), ("term_nil", C.c_uint8), ("term_ns_off", C.c_uint32), ("term_ns", C.c_int32),
              ("term_req_off", C.c_uint32), ("req_key", C.c_int32), ("req_op", C.c_int32), ("req_val_off", C.c_uint32),
              ("req_val", C.c_int32), ("pod_node", C.c_int32), ("pod_listed", C.c_uint8), ("pod_in_tasks", C.c_uint8),
              ("pod_unbound", C.c_uint8)]


class kbo_pod_objects(C.Structure):
    _fields_ = [("P", C.c_uint32), ("T", C.c_uint32)] + [(n, C.POINTER(t)) for n, t in _PO_FIELDS] + \
               [("n_topo", C.c_uint32), ("node_topo", C.POINTER(C.c_int32))]


def _serialize_pod_objects(raw) -> dict:
    """builder.flatten's raw pod objects (snap.meta["pod_objects"]) -> the interned arrays of kbo_pod_objects."""
    nodes, pending, existing, nidx = raw["nodes"], raw["pending"], raw["existing"], raw["node_index"]
    N, T = len(nodes), len(pending)
    strs = {}

    def sid(x: str) -> int:
        return strs.setdefault(x, len(strs))
    allpods = list(pending) + list(existing)
    topo_keys = {}
    ob = {"P": len(allpods), "T": T, "pod_ns": [], "lab_off": [0], "lab_key": [], "lab_val": [], "has_aff": [], "has_anti": [],
          "term_off": [0], "term_kind": [], "term_weight": [], "term_topo": [], "term_nil": [], "term_ns_off": [0], "term_ns": [],
          "term_req_off": [0], "req_key": [], "req_op": [], "req_val_off": [0], "req_val": [],
          "pod_node": [], "pod_listed": [], "pod_in_tasks": [], "pod_unbound": []}
    opn = {"In": 0, "NotIn": 1, "Exists": 2, "DoesNotExist": 3}
    for p in allpods:
        ob["pod_ns"].append(sid(p.namespace))
        for k, v in sorted(p.labels.items()):
            ob["lab_key"].append(sid(k)); ob["lab_val"].append(sid(v))
        ob["lab_off"].append(len(ob["lab_key"]))
        ob["has_aff"].append(1 if p.pod_affinity is not None else 0)
        ob["has_anti"].append(1 if p.pod_anti_affinity is not None else 0)
        for kind, a in ((0, p.pod_affinity), (1, p.pod_anti_affinity)):
            if a is None:
                continue
            for (pref, wt, term) in [(0, 0, t) for t in a.required] + [(2, w, t) for (w, t) in a.preferred]:
                ob["term_kind"].append(kind + pref)
                ob["term_weight"].append(wt)
                ob["term_topo"].append(topo_keys.setdefault(term.topology_key, len(topo_keys)) if term.topology_key else -1)
                ob["term_nil"].append(1 if term.nil_selector else 0)
                for ns in term.namespaces:
                    ob["term_ns"].append(sid(ns))
                ob["term_ns_off"].append(len(ob["term_ns"]))
                for k, v in sorted(term.match_labels.items()):      # matchLabels: key In (value) (LabelSelectorAsSelector)
                    ob["req_key"].append(sid(k)); ob["req_op"].append(0); ob["req_val"].append(sid(v)); ob["req_val_off"].append(len(ob["req_val"]))
                for (k, op, vals) in term.match_expressions:
                    ob["req_key"].append(sid(k)); ob["req_op"].append(opn[op])
                    for v in vals:
                        ob["req_val"].append(sid(v))
                    ob["req_val_off"].append(len(ob["req_val"]))
                ob["term_req_off"].append(len(ob["req_key"]))
        ob["term_off"].append(len(ob["term_kind"]))
    for p, li, it in zip(existing, raw["listed"], raw["in_tasks"]):
        ob["pod_node"].append(nidx[p.node_name])
        ob["pod_listed"].append(1 if li else 0)
        ob["pod_in_tasks"].append(1 if it else 0)
        ob["pod_unbound"].append(0)                                   # a pod the cache placed on a node has Spec.NodeName set
    ob["n_topo"] = len(topo_keys)
    nt = np.full((max(1, len(topo_keys)), max(1, N)), -1, dtype=np.int32)
    for key, ki in topo_keys.items():
        for i, n in enumerate(nodes):
            if key in n.labels:
                nt[ki, i] = sid("\0v:" + n.labels[key])
    ob["node_topo"] = nt
    return ob


class _PodObjects:
    """Hands the raw pod objects of a snapshot (builder.py: snap.meta["pod_objects"]) to the oracle for the duration of a call."""

    def __init__(self, snap: Snapshot):
        raw = (snap.meta or {}).get("pod_objects")
        if raw is not None and "P" not in raw:
            if "_kbo" not in raw:
                raw["_kbo"] = _serialize_pod_objects(raw)
            raw = raw["_kbo"]
        self.ob = raw
        self.N = snap.N

    def __enter__(self):
        if self.ob is None:
            return self
        ob = self.ob
        po = kbo_pod_objects()
        po.P, po.T, po.n_topo = int(ob["P"]), int(ob["T"]), int(ob["n_topo"])
        keep = []
        for n, t in _PO_FIELDS:
            a = np.ascontiguousarray(ob[n], dtype=np.dtype(t))
            if a.size == 0:
                a = np.zeros(1, dtype=np.dtype(t))
            keep.append(a)
            setattr(po, n, _p(a, t))
        nt = np.ascontiguousarray(ob["node_topo"], dtype=np.int32)
        keep.append(nt)
        po.node_topo = _p(nt, C.c_int32)
        lib().kbo_set_pod_objects(C.byref(po), C.c_uint32(self.N))
        return self

    def __exit__(self, *a):
        if self.ob is not None:
            lib().kbo_set_pod_objects(None, C.c_uint32(0))
        return False


def allocate(snap: Snapshot, conf: PluginConf, mode: int = KBO_MODE_OPTIMISED, threads: int = 1,
             max_tasks: int = 0, max_seconds: float = 0.0, actions: int = KBO_ACTION_ALLOCATE, warm_tasks: int = 0) -> OracleOut:
    """warm_tasks (timing samples): run that many tasks with cached aggregates first, then switch to `mode` and start the clock."""
    L = lib()
    cs, keep1 = snap.to_c()
    cc, keep2 = conf.to_c()
    o = kbo_opts(mode, threads, max_tasks, max_seconds, actions, warm_tasks)
    R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
    dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
    res = kbo_result()
    out = OracleOut(
        decisions=dec, result=res,
        node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
        node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
        node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
        job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
        queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
    with _PodObjects(snap):
        rc = L.kbo_allocate(C.byref(cs), C.byref(cc), C.byref(o), dec.ctypes.data_as(C.c_void_p), C.byref(res),
                            _p(out.node_idle, C.c_double), _p(out.node_releasing, C.c_double), _p(out.node_used, C.c_double),
                            _p(out.node_pods, C.c_int32), _p(out.node_nz_cpu, C.c_int64), _p(out.node_nz_mem, C.c_int64),
                            _p(out.node_ports, C.c_uint64), _p(out.job_share, C.c_double), _p(out.job_ready, C.c_int32),
                            _p(out.queue_share, C.c_double), _p(out.queue_deserved, C.c_double),
                            _p(out.queue_allocated, C.c_double))
    if rc != 0:
        raise RuntimeError(f"kbo_allocate rc={rc}: {L.kbo_last_error().decode()}")
    out.decisions = dec[:T]
    return out


class kbo_running(C.Structure):
    _fields_ = [("n", C.c_uint32), ("node", C.POINTER(C.c_uint32)), ("job", C.POINTER(C.c_uint32)), ("resreq", C.POINTER(C.c_double)),
                ("res_present", C.POINTER(C.c_uint32)), ("prio", C.POINTER(C.c_int32)), ("ctime", C.POINTER(C.c_int64)),
                ("uid_rank", C.POINTER(C.c_uint32)), ("flags", C.POINTER(C.c_uint32))]


ACTIONS = {"reclaim": 0, "allocate": 1, "backfill": 2, "preempt": 3}      # KBO_ACT_*


def cycle(snap: Snapshot, conf: PluginConf, actions=("allocate",), running: Optional[dict] = None, threads: int = 1):
    """One scheduling cycle on ONE session: `actions` in order (scheduler.go:88-101).  `running` = the table of Running tasks
    (builder.flatten() puts it into snap.meta["running"]); needed by reclaim / preempt.  Returns (OracleOut, evicted, evict_order)."""
    L = lib()
    cs, keep1 = snap.to_c()
    cc, keep2 = conf.to_c()
    o = kbo_opts(KBO_MODE_OPTIMISED, threads, 0, 0.0, 0, 0)
    R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
    dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
    res = kbo_result()
    out = OracleOut(
        decisions=dec, result=res,
        node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
        node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
        node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
        job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
        queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
    run = None
    nrun = 0
    keep3 = []
    if running is not None and len(running["node"]):
        nrun = len(running["node"])
        arrs = {k: np.ascontiguousarray(running[k], dtype=dt) for k, dt in
                (("node", np.uint32), ("job", np.uint32), ("resreq", np.float64), ("res_present", np.uint32), ("prio", np.int32),
                 ("ctime", np.int64), ("uid_rank", np.uint32), ("flags", np.uint32))}
        assert arrs["resreq"].shape == (R, nrun)
        keep3 = list(arrs.values())
        run = kbo_running(nrun, _p(arrs["node"], C.c_uint32), _p(arrs["job"], C.c_uint32), _p(arrs["resreq"], C.c_double),
                          _p(arrs["res_present"], C.c_uint32), _p(arrs["prio"], C.c_int32), _p(arrs["ctime"], C.c_int64),
                          _p(arrs["uid_rank"], C.c_uint32), _p(arrs["flags"], C.c_uint32))
    evicted = np.zeros(max(nrun, 1), dtype=np.uint8)
    order = np.zeros(max(nrun, 1), dtype=np.uint32)
    acts = np.array([ACTIONS[a] for a in actions], dtype=np.uint8)
    with _PodObjects(snap):
        rc = L.kbo_cycle(C.byref(cs), C.byref(run) if run is not None else None, C.byref(cc), C.byref(o),
                         _p(acts, C.c_uint8), C.c_uint32(len(acts)), dec.ctypes.data_as(C.c_void_p), _p(evicted, C.c_uint8), _p(order, C.c_uint32),
                         C.byref(res),
                         _p(out.node_idle, C.c_double), _p(out.node_releasing, C.c_double), _p(out.node_used, C.c_double),
                         _p(out.node_pods, C.c_int32), _p(out.node_nz_cpu, C.c_int64), _p(out.node_nz_mem, C.c_int64),
                         _p(out.node_ports, C.c_uint64), _p(out.job_share, C.c_double), _p(out.job_ready, C.c_int32),
                         _p(out.queue_share, C.c_double), _p(out.queue_deserved, C.c_double), _p(out.queue_allocated, C.c_double))
    if rc != 0:
        raise RuntimeError(f"kbo_cycle rc={rc}: {L.kbo_last_error().decode()}")
    out.decisions = dec[:T]
    del keep3
    return out, evicted[:nrun].astype(bool), order[:nrun]


def predicate_score(snap: Snapshot, conf: PluginConf, task: int):
    L = lib()
    cs, keep1 = snap.to_c()
    cc, keep2 = conf.to_c()
    fit = np.zeros(snap.N, dtype=np.uint8)
    score = np.zeros(snap.N, dtype=np.float64)
    with _PodObjects(snap):
        rc = L.kbo_predicate_score(C.byref(cs), C.byref(cc), C.c_uint32(task), _p(fit, C.c_uint8), _p(score, C.c_double))
    if rc != 0:
        raise RuntimeError(f"kbo_predicate_score rc={rc}: {L.kbo_last_error().decode()}")
    return fit, score


# ---- unit-level helpers (golden-vector tests) ----
def _res(v, R):
    a = np.zeros(R, dtype=np.float64)
    a[: len(v)] = v
    return a


def res_less_equal(l, lp, r, rp, R=3) -> bool:
    a, b = _res(l, R), _res(r, R)
    return bool(lib().kbo_res_less_equal(R, _p(a, C.c_double), C.c_uint32(lp), _p(b, C.c_double), C.c_uint32(rp)))


def res_less(l, lp, r, rp, R=3) -> bool:
    a, b = _res(l, R), _res(r, R)
    return bool(lib().kbo_res_less(R, _p(a, C.c_double), C.c_uint32(lp), _p(b, C.c_double), C.c_uint32(rp)))


def res_is_empty(l, lp, R=3) -> bool:
    a = _res(l, R)
    return bool(lib().kbo_res_is_empty(R, _p(a, C.c_double), C.c_uint32(lp)))


def _binop(fn, l, lp, r, rp, R):
    a, b = _res(l, R), _res(r, R)
    p = C.c_uint32(lp)
    rc = fn(R, _p(a, C.c_double), C.byref(p), _p(b, C.c_double), C.c_uint32(rp))
    return a, p.value, rc


def res_add(l, lp, r, rp, R=3):
    a, p, _ = _binop(lib().kbo_res_add, l, lp, r, rp, R)
    return a, p


def res_sub(l, lp, r, rp, R=3):
    a, p, rc = _binop(lib().kbo_res_sub, l, lp, r, rp, R)
    if rc != 0:
        raise ArithmeticError("Resource is not sufficient to do operation")
    return a, p


def res_set_max(l, lp, r, rp, R=3):
    a, p, _ = _binop(lib().kbo_res_set_max, l, lp, r, rp, R)
    return a, p


def res_fit_delta(l, lp, r, rp, R=3):
    a, p, _ = _binop(lib().kbo_res_fit_delta, l, lp, r, rp, R)
    return a, p


def heap_sort(keys):
    k = np.asarray(keys, dtype=np.int64)
    out = np.zeros_like(k)
    lib().kbo_heap_sort(_p(k, C.c_int64), C.c_uint32(len(k)), _p(out, C.c_int64))
    return out
