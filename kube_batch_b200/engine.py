"""ctypes binding of libkbgpu.so — the product path.  There is no fallback: if the shared object
or a CUDA device is missing, construction raises."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .snapshot import PluginConf, Snapshot

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkbgpu.so")

EXPORTS = [
    "kb_engine_create", "kb_engine_destroy", "kb_session_load", "kb_allocate", "kb_backfill", "kb_predicate_score",
    "kb_best_nodes", "kb_node_state", "kb_order_state", "kb_last_error", "kb_status_str", "kb_version",
    "kb_nccl_unique_id", "kb_last_kernel_ms", "kb_session_load_running", "kb_reclaim", "kb_preempt", "kb_cycle", "kb_bind_list",
]


class KbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{msg} (status {code})")
        self.code = code


_lib = None


def load_library():
    """dlopen libkbgpu.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.kb_last_error.restype = C.c_char_p
        L.kb_last_error.argtypes = [C.c_void_p]
        L.kb_status_str.restype = C.c_char_p
        L.kb_version.restype = C.c_char_p
        L.kb_engine_destroy.restype = None
        L.kb_engine_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


@dataclass
class CycleResult:
    decisions: np.ndarray          # structured, abi.DECISION_DTYPE, one per snapshot task
    stats: abi.kb_stats

    def bind_map(self):
        d = self.decisions
        return {int(t): int(d["node"][t]) for t in np.nonzero(d["dispatched"])[0]}


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Engine:
    """One kb_engine.  Not thread-safe (like the single runOnce goroutine of the reference)."""

    def __init__(self, device: int = 0, rank: int = 0, world_size: int = 1, nccl_unique_id: Optional[bytes] = None, flags: int = 0):
        self.L = load_library()
        self._h = C.c_void_p()
        opts = abi.kb_engine_opts()
        opts.abi_version = abi.KB_ABI_VERSION
        opts.device = device
        opts.rank = rank
        opts.world_size = world_size
        self._uid = (C.c_char * 128).from_buffer_copy(nccl_unique_id) if nccl_unique_id else None
        opts.nccl_unique_id = C.cast(self._uid, C.c_void_p) if self._uid is not None else None
        opts.flags = flags
        rc = self.L.kb_engine_create(C.byref(opts), C.byref(self._h))
        if rc != 0:
            raise KbError(rc, "kb_engine_create: " + self.L.kb_last_error(None).decode())
        self.snap: Optional[Snapshot] = None

    def close(self):
        if self._h:
            self.L.kb_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise KbError(rc, f"{what}: {self.L.kb_last_error(self._h).decode()}")

    def load(self, snap: Snapshot, conf: PluginConf):
        cs, keep1 = snap.to_c()
        cc, keep2 = conf.to_c()
        self._check(self.L.kb_session_load(self._h, C.byref(cs), C.byref(cc)), "kb_session_load")
        self.snap = snap
        return self

    def allocate(self) -> CycleResult:
        T = self.snap.T
        dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
        st = abi.kb_stats()
        self._check(self.L.kb_allocate(self._h, dec.ctypes.data_as(C.c_void_p), C.byref(st)), "kb_allocate")
        return CycleResult(dec[:T], st)

    def backfill(self) -> CycleResult:
        """backfillAction.Execute (actions/backfill/backfill.go:40-71) on the current device state: call it after
        allocate() for the default action list "allocate, backfill".  Returns the full, updated decision table."""
        T = self.snap.T
        dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
        st = abi.kb_stats()
        self._check(self.L.kb_backfill(self._h, dec.ctypes.data_as(C.c_void_p), C.byref(st)), "kb_backfill")
        return CycleResult(dec[:T], st)

    def load_running(self, running: Optional[dict]):
        """Hands the Running tasks (builder.py: snapshot.meta["running"]) to the engine: needed by reclaim() / preempt()."""
        cs, keep1 = self.snap.to_c()
        cr, keep2 = abi.running_to_c(running, self.snap.R, self.snap.J)
        self._check(self.L.kb_session_load_running(self._h, C.byref(cs), C.byref(cr)), "kb_session_load_running")
        self._n_run = int(cr.n)
        return self

    def _evict(self, fn, what):
        T, n = self.snap.T, self._n_run
        dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
        ev = np.zeros(max(n, 1), dtype=np.uint8)
        order = np.zeros(max(n, 1), dtype=np.uint32)
        st = abi.kb_stats()
        self._check(fn(self._h, dec.ctypes.data_as(C.c_void_p), _p(ev, C.c_uint8), _p(order, C.c_uint32), C.byref(st)), what)
        return CycleResult(dec[:T], st), ev[:n].astype(bool), order[:n]

    def reclaim(self):
        """reclaimAction.Execute (actions/reclaim/reclaim.go:41-193) from the loaded state -> (CycleResult, evicted, evict_order)."""
        return self._evict(self.L.kb_reclaim, "kb_reclaim")

    def preempt(self):
        """preemptAction.Execute (actions/preempt/preempt.go:43-270) from the loaded state -> (CycleResult, evicted, evict_order)."""
        return self._evict(self.L.kb_preempt, "kb_preempt")

    ACTIONS = {"reclaim": 0, "allocate": 1, "backfill": 2, "preempt": 3}      # KB_ACT_*

    def cycle(self, actions=("reclaim", "allocate", "backfill", "preempt")):
        """One scheduling cycle on ONE session (scheduler.go:88-101): the action list from the loaded state.
        -> (CycleResult, evicted, evict_order, bounds[n_actions][2])."""
        T, n = self.snap.T, getattr(self, "_n_run", 0)
        acts = np.array([self.ACTIONS[a] for a in actions], dtype=np.uint8)
        dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
        ev = np.zeros(max(n, 1), dtype=np.uint8)
        order = np.zeros(max(n, 1), dtype=np.uint32)
        bounds = np.zeros((max(len(acts), 1), 2), dtype=np.uint32)
        st = abi.kb_stats()
        self._check(self.L.kb_cycle(self._h, _p(acts, C.c_uint8), C.c_uint32(len(acts)), dec.ctypes.data_as(C.c_void_p), _p(ev, C.c_uint8),
                                    _p(order, C.c_uint32), _p(bounds, C.c_uint32), C.byref(st)), "kb_cycle")
        return CycleResult(dec[:T], st), ev[:n].astype(bool), order[:n], bounds[: len(acts)]

    def bind_list(self):
        """The (task, node) pairs that reach cache.Bind, in ssn.dispatch order (device-side compaction + radix sort)."""
        T = self.snap.T
        task = np.zeros(max(T, 1), dtype=np.uint32)
        node = np.zeros(max(T, 1), dtype=np.int32)
        n = C.c_uint32(0)
        self._check(self.L.kb_bind_list(self._h, _p(task, C.c_uint32), _p(node, C.c_int32), C.byref(n)), "kb_bind_list")
        return task[: n.value], node[: n.value]

    def predicate_score(self, lo: int, hi: int, want_score: bool = True):
        """kb_predicate_score for tasks [lo, hi) against the current device state.  want_score=False: fit only (sessions whose
        priorities need a reduction over the feasible nodes — inter-pod / preferred node affinity on the counter path)."""
        N = self.snap.N
        fit = np.zeros((hi - lo, N), dtype=np.uint8)
        score = np.zeros((hi - lo, N), dtype=np.float64) if want_score else None
        self._check(self.L.kb_predicate_score(self._h, C.c_uint32(lo), C.c_uint32(hi), _p(fit, C.c_uint8),
                                              _p(score, C.c_double) if want_score else None), "kb_predicate_score")
        return fit, score

    def best_nodes(self, lo: int, hi: int) -> np.ndarray:
        out = np.zeros(max(hi - lo, 1), dtype=np.uint64)
        self._check(self.L.kb_best_nodes(self._h, C.c_uint32(lo), C.c_uint32(hi), _p(out, C.c_uint64)), "kb_best_nodes")
        return out[: hi - lo]

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self.L.kb_last_kernel_ms(self._h, C.byref(ms)), "kb_last_kernel_ms")
        return float(ms.value)

    def node_state(self):
        s = self.snap
        out = dict(idle=np.zeros((s.R, s.N)), releasing=np.zeros((s.R, s.N)), used=np.zeros((s.R, s.N)),
                   pods=np.zeros(s.N, dtype=np.int32), nz_cpu=np.zeros(s.N, dtype=np.int64),
                   nz_mem=np.zeros(s.N, dtype=np.int64), ports=np.zeros((s.W, s.N), dtype=np.uint64))
        self._check(self.L.kb_node_state(self._h, _p(out["idle"], C.c_double), _p(out["releasing"], C.c_double),
                                         _p(out["used"], C.c_double), _p(out["pods"], C.c_int32),
                                         _p(out["nz_cpu"], C.c_int64), _p(out["nz_mem"], C.c_int64),
                                         _p(out["ports"], C.c_uint64)), "kb_node_state")
        return out

    def order_state(self):
        s = self.snap
        out = dict(job_share=np.zeros(s.J), job_ready=np.zeros(s.J, dtype=np.int32), queue_share=np.zeros(s.Q),
                   queue_deserved=np.zeros((s.R, s.Q)), queue_allocated=np.zeros((s.R, s.Q)))
        self._check(self.L.kb_order_state(self._h, _p(out["job_share"], C.c_double), _p(out["job_ready"], C.c_int32),
                                          _p(out["queue_share"], C.c_double), _p(out["queue_deserved"], C.c_double),
                                          _p(out["queue_allocated"], C.c_double)), "kb_order_state")
        return out


def nccl_unique_id() -> bytes:
    """A fresh ncclUniqueId (rank 0); broadcast it and pass it to every rank's Engine(...)."""
    L = load_library()
    buf = (C.c_char * 128)()
    rc = L.kb_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise KbError(rc, "kb_nccl_unique_id: " + L.kb_last_error(None).decode())
    return bytes(buf.raw)


def key_node(key: np.ndarray) -> np.ndarray:
    """Decode the packed best key of kb_best_nodes: node index, -1 where no node fits."""
    k = np.asarray(key, dtype=np.uint64)
    node = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    return np.where(k == 0, -1, node)
