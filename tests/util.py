"""Shared test helpers: CPU emulation binding + result comparison against the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from kube_batch_b200 import abi
from kube_batch_b200.snapshot import PluginConf, Snapshot
from oracle import kbo

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = os.path.join(_HERE, "emu", "libkbemu.so")
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        import fcntl
        with open(os.path.join(_HERE, "emu", ".build.lock"), "w") as lk:     # ranks / workers build one at a time
            fcntl.flock(lk, fcntl.LOCK_EX)
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "emu"), "-s"])
        _emu = C.CDLL(_EMU)
        _emu.kbemu_last_error.restype = C.c_char_p
    return _emu


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def emu_allocate(snap: Snapshot, conf: PluginConf, actions: int = 1, mode: int = 0) -> kbo.OracleOut:
    """Run the CPU emulation of the device algorithm; same result container as the oracle.
    actions: bit 0 = allocate (kb_allocate), bit 1 = backfill afterwards (kb_backfill).
    mode: 0 = scan/replay overlap protocol, 1 = plain one-class launches, 2 / 4 = chained visits (visit_chain_kernel<K>)."""
    L = emu_lib()
    cs, k1 = snap.to_c()
    cc, k2 = conf.to_c()
    R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
    dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
    st = abi.kb_stats()
    out = kbo.OracleOut(
        decisions=dec, result=st,
        node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
        node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
        node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
        job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
        queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
    rc = L.kbemu_allocate(C.byref(cs), C.byref(cc), C.c_uint32(actions), C.c_uint32(mode), dec.ctypes.data_as(C.c_void_p), C.byref(st),
                          _p(out.node_idle, C.c_double), _p(out.node_releasing, C.c_double), _p(out.node_used, C.c_double),
                          _p(out.node_pods, C.c_int32), _p(out.node_nz_cpu, C.c_int64), _p(out.node_nz_mem, C.c_int64),
                          _p(out.node_ports, C.c_uint64), _p(out.job_share, C.c_double), _p(out.job_ready, C.c_int32),
                          _p(out.queue_share, C.c_double), _p(out.queue_deserved, C.c_double), _p(out.queue_allocated, C.c_double))
    if rc != 0:
        raise RuntimeError(f"kbemu_allocate rc={rc}: {L.kbemu_last_error().decode()}")
    out.decisions = dec[:T]
    return out


def emu_evict(snap: Snapshot, conf: PluginConf, action: str, running):
    """kb_evict.h (reclaim / preempt) run by the CPU emulation from the as-loaded state -> (OracleOut, evicted, evict_order)."""
    L = emu_lib()
    cs, k1 = snap.to_c()
    cc, k2 = conf.to_c()
    cr, k3 = abi.running_to_c(running, snap.R, snap.J)
    R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
    n = int(cr.n)
    dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
    ev = np.zeros(max(n, 1), dtype=np.uint8)
    order = np.zeros(max(n, 1), dtype=np.uint32)
    st = abi.kb_stats()
    out = kbo.OracleOut(
        decisions=dec, result=st,
        node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
        node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
        node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
        job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
        queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
    rc = L.kbemu_evict(C.byref(cs), C.byref(cr), C.byref(cc), C.c_uint32({"reclaim": 0, "preempt": 1}[action]),
                       dec.ctypes.data_as(C.c_void_p), _p(ev, C.c_uint8), _p(order, C.c_uint32), C.byref(st),
                       _p(out.node_idle, C.c_double), _p(out.node_releasing, C.c_double), _p(out.node_used, C.c_double),
                       _p(out.node_pods, C.c_int32), _p(out.node_nz_cpu, C.c_int64), _p(out.node_nz_mem, C.c_int64),
                       _p(out.node_ports, C.c_uint64), _p(out.job_share, C.c_double), _p(out.job_ready, C.c_int32),
                       _p(out.queue_share, C.c_double), _p(out.queue_deserved, C.c_double), _p(out.queue_allocated, C.c_double))
    if rc != 0:
        raise RuntimeError(f"kbemu_evict rc={rc}: {L.kbemu_last_error().decode()}")
    out.decisions = dec[:T]
    return out, ev[:n].astype(bool), order[:n]


def emu_cycle(snap: Snapshot, conf: PluginConf, actions, running, mode: int = 1):
    """kb_cycle on the CPU emulation: the action list on ONE session -> (OracleOut, evicted, evict_order)."""
    L = emu_lib()
    cs, k1 = snap.to_c()
    cc, k2 = conf.to_c()
    cr, k3 = abi.running_to_c(running, snap.R, snap.J)
    R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
    n = int(cr.n)
    dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
    ev = np.zeros(max(n, 1), dtype=np.uint8)
    order = np.zeros(max(n, 1), dtype=np.uint32)
    st = abi.kb_stats()
    acts = np.array([kbo.ACTIONS[a] for a in actions], dtype=np.uint8)
    out = kbo.OracleOut(
        decisions=dec, result=st,
        node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
        node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
        node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
        job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
        queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
    rc = L.kbemu_cycle(C.byref(cs), C.byref(cr), C.byref(cc), _p(acts, C.c_uint8), C.c_uint32(len(acts)), C.c_uint32(mode),
                       dec.ctypes.data_as(C.c_void_p), _p(ev, C.c_uint8), _p(order, C.c_uint32), C.byref(st),
                       _p(out.node_idle, C.c_double), _p(out.node_releasing, C.c_double), _p(out.node_used, C.c_double),
                       _p(out.node_pods, C.c_int32), _p(out.node_nz_cpu, C.c_int64), _p(out.node_nz_mem, C.c_int64),
                       _p(out.node_ports, C.c_uint64), _p(out.job_share, C.c_double), _p(out.job_ready, C.c_int32),
                       _p(out.queue_share, C.c_double), _p(out.queue_deserved, C.c_double), _p(out.queue_allocated, C.c_double))
    if rc != 0:
        raise RuntimeError(f"kbemu_cycle rc={rc}: {L.kbemu_last_error().decode()}")
    out.decisions = dec[:T]
    return out, ev[:n].astype(bool), order[:n]


def assert_same_decisions(ref: np.ndarray, got: np.ndarray, what: str = ""):
    """Bit-exact placement parity: node, kind, dispatched, step and dispatch_step of every task."""
    assert ref.shape == got.shape, (what, ref.shape, got.shape)
    for f in ("kind", "node", "step", "dispatched", "dispatch_step"):
        bad = np.nonzero(ref[f] != got[f])[0]
        if len(bad):
            t = int(bad[0])
            raise AssertionError(f"{what}: field '{f}' differs for {len(bad)} task(s); first task {t}: "
                                 f"oracle={ref[t]} got={got[t]}")


def assert_same_state(o: kbo.OracleOut, node_state: dict, order_state: dict, what: str = ""):
    np.testing.assert_array_equal(o.node_idle, node_state["idle"], err_msg=f"{what}: node idle")
    np.testing.assert_array_equal(o.node_releasing, node_state["releasing"], err_msg=f"{what}: node releasing")
    np.testing.assert_array_equal(o.node_used, node_state["used"], err_msg=f"{what}: node used")
    np.testing.assert_array_equal(o.node_pods, node_state["pods"], err_msg=f"{what}: pods")
    np.testing.assert_array_equal(o.node_nz_cpu, node_state["nz_cpu"], err_msg=f"{what}: nz_cpu")
    np.testing.assert_array_equal(o.node_nz_mem, node_state["nz_mem"], err_msg=f"{what}: nz_mem")
    np.testing.assert_array_equal(o.node_ports, node_state["ports"], err_msg=f"{what}: ports")
    np.testing.assert_array_equal(o.job_share, order_state["job_share"], err_msg=f"{what}: job share")
    np.testing.assert_array_equal(o.job_ready, order_state["job_ready"], err_msg=f"{what}: job ready")
    np.testing.assert_array_equal(o.queue_share, order_state["queue_share"], err_msg=f"{what}: queue share")
    np.testing.assert_array_equal(o.queue_deserved, order_state["queue_deserved"], err_msg=f"{what}: deserved")
    np.testing.assert_array_equal(o.queue_allocated, order_state["queue_allocated"], err_msg=f"{what}: allocated")


def emu_states(e: kbo.OracleOut):
    ns = dict(idle=e.node_idle, releasing=e.node_releasing, used=e.node_used, pods=e.node_pods, nz_cpu=e.node_nz_cpu,
              nz_mem=e.node_nz_mem, ports=e.node_ports)
    os_ = dict(job_share=e.job_share, job_ready=e.job_ready, queue_share=e.queue_share, queue_deserved=e.queue_deserved,
               queue_allocated=e.queue_allocated)
    return ns, os_


# ------------------------------------------------------------------------------------------------
# sharded node axis on CPU: one process per rank, torch.distributed / gloo instead of NCCL
# ------------------------------------------------------------------------------------------------
def emu_sharded_rank(rank: int, world: int, port: int, snap: Snapshot, conf: PluginConf, actions: int = 1):
    """Run the scan -> all-gather -> replay chain of the multi-GPU path with the CPU emulation and gloo.
    Returns the same container as emu_allocate (every rank holds the full, identical result)."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = emu_lib()
        L.kbemu_create.restype = C.c_void_p
        L.kbemu_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.kbemu_buf_u64.argtypes = [C.c_void_p]
        L.kbemu_buf_u64.restype = C.c_uint32
        for f in ("kbemu_scan", "kbemu_replay"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, f).restype = None
        L.kbemu_done.argtypes = [C.c_void_p]
        L.kbemu_destroy.argtypes = [C.c_void_p]
        cs, k1 = snap.to_c()
        cc, k2 = conf.to_c()
        h = L.kbemu_create(C.addressof(cs), C.addressof(cc), rank, world)
        if not h:
            raise RuntimeError(L.kbemu_last_error().decode())
        n = L.kbemu_buf_u64(h)
        send = torch.zeros(n, dtype=torch.int64)
        recv = torch.zeros(world * n, dtype=torch.int64)
        L.kbemu_begin_backfill.argtypes = [C.c_void_p, C.c_int]
        L.kbemu_begin_backfill.restype = None
        guard = 0
        for bit in (1, 2):
            if not (actions & bit):
                continue
            if bit == 2:
                L.kbemu_begin_backfill(h, 1 if (actions & 1) else 0)                        # kb_backfill: switch to the backfill view
            while not L.kbemu_done(h):
                L.kbemu_scan(h, C.c_void_p(send.data_ptr()))
                dist.all_gather_into_tensor(recv, send)          # == ncclAllGather(sendbuf, recvbuf, ...)
                L.kbemu_replay(h, C.c_void_p(recv.data_ptr()))
                guard += 1
                assert guard < 10_000_000
        R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
        dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
        st = abi.kb_stats()
        out = kbo.OracleOut(
            decisions=dec, result=st,
            node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
            node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
            node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
            job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
            queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
        L.kbemu_finish.argtypes = [C.c_void_p] + [C.c_void_p] * 14
        L.kbemu_finish(h, dec.ctypes.data, C.addressof(st), out.node_idle.ctypes.data, out.node_releasing.ctypes.data,
                       out.node_used.ctypes.data, out.node_pods.ctypes.data, out.node_nz_cpu.ctypes.data,
                       out.node_nz_mem.ctypes.data, out.node_ports.ctypes.data, out.job_share.ctypes.data,
                       out.job_ready.ctypes.data, out.queue_share.ctypes.data, out.queue_deserved.ctypes.data,
                       out.queue_allocated.ctypes.data)
        out.decisions = dec[:T]
        L.kbemu_destroy(h)
        return out
    finally:
        dist.destroy_process_group()


def emu_sharded_inprocess(snap: Snapshot, conf: PluginConf, world: int, actions: int = 1, mode: int = 0):
    """The sharded scan -> all-gather -> replay chain with all `world` ranks emulated in THIS process (the all-gather is a
    concatenation).  Returns one result container per rank; every rank must hold the identical, complete result."""
    L = emu_lib()
    L.kbemu_create.restype = C.c_void_p
    L.kbemu_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.kbemu_buf_u64.argtypes = [C.c_void_p]
    L.kbemu_buf_u64.restype = C.c_uint32
    for f in ("kbemu_scan", "kbemu_replay"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        getattr(L, f).restype = None
    L.kbemu_done.argtypes = [C.c_void_p]
    L.kbemu_destroy.argtypes = [C.c_void_p]
    L.kbemu_begin_backfill.argtypes = [C.c_void_p, C.c_int]
    L.kbemu_begin_backfill.restype = None
    L.kbemu_finish.argtypes = [C.c_void_p] + [C.c_void_p] * 14
    cs, k1 = snap.to_c()
    cc, k2 = conf.to_c()
    hs = []
    try:
        L.kbemu_create2.restype = C.c_void_p
        L.kbemu_create2.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        for r in range(world):
            h = L.kbemu_create2(C.addressof(cs), C.addressof(cc), r, world, mode)
            if not h:
                raise RuntimeError(L.kbemu_last_error().decode())
            hs.append(h)
        n = L.kbemu_buf_u64(hs[0])
        send = [np.zeros(n, dtype=np.uint64) for _ in range(world)]
        guard = 0
        for bit in (1, 2):
            if not (actions & bit):
                continue
            if bit == 2:
                for h in hs:
                    L.kbemu_begin_backfill(h, 1 if (actions & 1) else 0)
            while not L.kbemu_done(hs[0]):
                for r, h in enumerate(hs):
                    L.kbemu_scan(h, send[r].ctypes.data)
                recv = np.concatenate(send)                  # == ncclAllGather / the peer-memory exchange
                for h in hs:
                    L.kbemu_replay(h, recv.ctypes.data)
                guard += 1
                assert guard < 10_000_000
            assert all(L.kbemu_done(h) for h in hs), "ranks disagree on the end of the cycle"
        outs = []
        R, W, N, T, J, Q = snap.R, snap.W, snap.N, snap.T, snap.J, snap.Q
        for h in hs:
            dec = np.zeros(max(T, 1), dtype=np.dtype(abi.DECISION_DTYPE))
            st = abi.kb_stats()
            out = kbo.OracleOut(
                decisions=dec, result=st,
                node_idle=np.zeros((R, N)), node_releasing=np.zeros((R, N)), node_used=np.zeros((R, N)),
                node_pods=np.zeros(N, dtype=np.int32), node_nz_cpu=np.zeros(N, dtype=np.int64),
                node_nz_mem=np.zeros(N, dtype=np.int64), node_ports=np.zeros((W, N), dtype=np.uint64),
                job_share=np.zeros(J), job_ready=np.zeros(J, dtype=np.int32), queue_share=np.zeros(Q),
                queue_deserved=np.zeros((R, Q)), queue_allocated=np.zeros((R, Q)))
            L.kbemu_finish(h, dec.ctypes.data, C.addressof(st), out.node_idle.ctypes.data, out.node_releasing.ctypes.data,
                           out.node_used.ctypes.data, out.node_pods.ctypes.data, out.node_nz_cpu.ctypes.data,
                           out.node_nz_mem.ctypes.data, out.node_ports.ctypes.data, out.job_share.ctypes.data,
                           out.job_ready.ctypes.data, out.queue_share.ctypes.data, out.queue_deserved.ctypes.data,
                           out.queue_allocated.ctypes.data)
            out.decisions = dec[:T]
            outs.append(out)
        return outs
    finally:
        for h in hs:
            L.kbemu_destroy(h)
