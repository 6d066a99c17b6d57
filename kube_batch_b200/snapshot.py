"""Flattened Session snapshot + plugin configuration (host side of the C ABI).

`Snapshot` is the SoA interchange format of include/kbgpu.h `kb_snapshot`: what the Go shim
 produces from ssn.Jobs / ssn.Nodes / ssn.Queues
(/root/reference/pkg/scheduler/framework/session.go:37-46), what libkbgpu.so consumes and what
the CPU oracle replays.  `PluginConf` mirrors conf.Tier / conf.PluginOption
(/root/reference/pkg/scheduler/conf/scheduler_conf.go:20-56).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import abi


def _shape(code: str, d: Dict[str, int]):
    out = []
    i = 0
    while i < len(code):
        ch = code[i]
        if ch == "J" and i + 1 < len(code) and code[i + 1] == "1":
            out.append(d["J"] + 1)
            i += 2
            continue
        out.append(d[ch])
        i += 1
    return tuple(out)


class Snapshot:
    """Dense SoA snapshot.  Attributes are numpy arrays named like the kb_snapshot fields."""

    def __init__(self, R: int, W: int, N: int, T: int, J: int, Q: int):
        assert 2 <= R <= abi.KB_MAX_R and 1 <= W <= abi.KB_MAX_W
        self.R, self.W, self.N, self.T, self.J, self.Q = R, W, N, T, J, Q
        dims = self.dims()
        for name, _, dt, shp in abi.SNAPSHOT_ARRAYS:
            setattr(self, name, np.zeros(_shape(shp, dims), dtype=np.dtype(dt)))
        self.meta: Dict[str, object] = {}
        self.pod_affinity = None       # dict mirroring kb_pod_affinity (builder.py), None = no pod carries inter-pod terms

    def dims(self) -> Dict[str, int]:
        return {"R": self.R, "W": self.W, "N": self.N, "T": self.T, "J": self.J, "Q": self.Q, "A": abi.KB_MAX_AFF_TERMS,
                "P": abi.KB_MAX_PREF_TERMS}

    def validate(self) -> None:
        dims = self.dims()
        for name, _, dt, shp in abi.SNAPSHOT_ARRAYS:
            a = getattr(self, name)
            assert a.dtype == np.dtype(dt), (name, a.dtype, dt)
            assert a.shape == _shape(shp, dims), (name, a.shape, _shape(shp, dims))
            assert a.flags["C_CONTIGUOUS"], name
        assert self.job_task_off[0] == 0 and self.job_task_off[-1] == self.T
        assert np.all(np.diff(self.job_task_off.astype(np.int64)) >= 0)
        assert np.all(self.task_resreq <= self.task_initreq)
        if self.T:
            assert len(np.unique(self.task_uid_rank)) == self.T

    def to_c(self):
        """Return (kb_snapshot, keepalive). The struct points into this object's arrays.
        Validation runs once per object (call `invalidate()` after editing arrays in place)."""
        if not getattr(self, "_validated", False):
            self.validate()
            self._validated = True
        s = abi.kb_snapshot()
        s.abi_version = abi.KB_ABI_VERSION
        s.R, s.W, s.N, s.T, s.J, s.Q = self.R, self.W, self.N, self.T, self.J, self.Q
        s.flags = int(getattr(self, "flags", 0))      # KB_SNAPSHOT_*
        keep = []
        for name, ptr, _, _ in abi.SNAPSHOT_ARRAYS:
            a = getattr(self, name)
            if a.size == 0:  # ctypes cannot take the address of an empty array
                a = np.zeros(1, dtype=a.dtype)
            keep.append(a)
            setattr(s, name, a.ctypes.data_as(ptr))
        if getattr(self, "pod_affinity", None) is not None:
            pa, k2 = abi.pod_affinity_to_c(self.pod_affinity)
            keep += [pa, k2]
            s.pod_affinity = C.pointer(pa)
        return s, keep

    def invalidate(self) -> None:
        self._validated = False

    def job_of_task(self) -> np.ndarray:
        j = np.zeros(self.T, dtype=np.int64)
        for k in range(self.J):
            j[self.job_task_off[k]:self.job_task_off[k + 1]] = k
        return j

    def save(self, path: str) -> None:
        arrs = {name: getattr(self, name) for name, *_ in abi.SNAPSHOT_ARRAYS}
        np.savez_compressed(path, __dims=np.array([self.R, self.W, self.N, self.T, self.J, self.Q]), **arrs)

    @staticmethod
    def load(path: str) -> "Snapshot":
        z = np.load(path)
        R, W, N, T, J, Q = [int(x) for x in z["__dims"]]
        s = Snapshot(R, W, N, T, J, Q)
        for name, *_ in abi.SNAPSHOT_ARRAYS:
            setattr(s, name, np.ascontiguousarray(z[name]))
        return s


# ----------------------------------------------------------------------------------------------
# plugin configuration
# ----------------------------------------------------------------------------------------------
_ENABLE_FIELDS = [
    "enabled_job_order", "enabled_job_ready", "enabled_job_pipelined", "enabled_task_order",
    "enabled_preemptable", "enabled_reclaimable", "enabled_queue_order", "enabled_predicate",
    "enabled_node_order",
]


@dataclass
class PluginOption:
    """conf.PluginOption (scheduler_conf.go:33-56).  None = the Go nil *bool."""
    name: str
    enabled_job_order: Optional[bool] = None
    enabled_job_ready: Optional[bool] = None
    enabled_job_pipelined: Optional[bool] = None
    enabled_task_order: Optional[bool] = None
    enabled_preemptable: Optional[bool] = None
    enabled_reclaimable: Optional[bool] = None
    enabled_queue_order: Optional[bool] = None
    enabled_predicate: Optional[bool] = None
    enabled_node_order: Optional[bool] = None
    arguments: Dict[str, str] = field(default_factory=dict)

    def apply_defaults(self) -> "PluginOption":
        """plugins.ApplyPluginConfDefaults (plugins/defaults.go:22-52): every nil Enabled* -> true."""
        for f in _ENABLE_FIELDS:
            if getattr(self, f) is None:
                setattr(self, f, True)
        return self


class PluginConf:
    """[]conf.Tier.  `tiers` is a list of lists of PluginOption."""

    def __init__(self, tiers: List[List[PluginOption]]):
        self.tiers = tiers

    @staticmethod
    def from_names(tiers: List[List[str]], arguments: Optional[Dict[str, Dict[str, str]]] = None) -> "PluginConf":
        """Like loadSchedulerConf (pkg/scheduler/util.go:44-73): names only, defaults applied."""
        arguments = arguments or {}
        return PluginConf([[PluginOption(n, arguments=dict(arguments.get(n, {}))).apply_defaults() for n in t] for t in tiers])

    @staticmethod
    def default() -> "PluginConf":
        """defaultSchedulerConf (pkg/scheduler/util.go:31-42)."""
        return PluginConf.from_names([["priority", "gang"], ["drf", "predicates", "proportion", "nodeorder"]])

    def to_c(self):
        keep = []
        ctiers = (abi.kb_tier * max(1, len(self.tiers)))()
        for ti, tier in enumerate(self.tiers):
            opts = (abi.kb_plugin_option * max(1, len(tier)))()
            for pi, p in enumerate(tier):
                o = opts[pi]
                o.name = p.name.encode()
                for f in _ENABLE_FIELDS:
                    setattr(o, f, 1 if getattr(p, f) else 0)
                keys = [k.encode() for k in p.arguments]
                vals = [str(v).encode() for v in p.arguments.values()]
                o.n_args = len(keys)
                ka = (C.c_char_p * max(1, len(keys)))(*keys)
                va = (C.c_char_p * max(1, len(vals)))(*vals)
                o.arg_keys = C.cast(ka, C.POINTER(C.c_char_p))
                o.arg_values = C.cast(va, C.POINTER(C.c_char_p))
                keep += [ka, va, keys, vals]
            ctiers[ti].n_plugins = len(tier)
            ctiers[ti].plugins = C.cast(opts, C.POINTER(abi.kb_plugin_option))
            keep.append(opts)
        c = abi.kb_plugin_conf()
        c.n_tiers = len(self.tiers)
        c.tiers = C.cast(ctiers, C.POINTER(abi.kb_tier))
        keep.append(ctiers)
        return c, keep

    def describe(self) -> str:
        return " | ".join(",".join(p.name for p in t) for t in self.tiers)
