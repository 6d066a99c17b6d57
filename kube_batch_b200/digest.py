"""Digest of a cycle's outcome: SHA-256 over the decision table (field by field, little endian) and the node / job
bookkeeping the cycle leaves behind.  bench.py compares the digest of the engine's cycle on EVERY rank with the committed
oracle digests (tests/golden/cycle_hashes.json, written by tests/golden/make_golden.py) so that every bench / scaling line
carries a parity verdict; the tests use the same function."""
from __future__ import annotations

import hashlib

import numpy as np

FIELDS = ("kind", "node", "step", "dispatched", "dispatch_step")


def decisions_digest(dec: np.ndarray) -> str:
    h = hashlib.sha256()
    for f in FIELDS:
        h.update(np.ascontiguousarray(dec[f]).tobytes())
    return h.hexdigest()


def state_digest(node_idle: np.ndarray, node_releasing: np.ndarray, job_ready: np.ndarray, job_share: np.ndarray) -> str:
    h = hashlib.sha256()
    for a, t in ((node_idle, np.float64), (node_releasing, np.float64), (job_ready, np.int32), (job_share, np.float64)):
        h.update(np.ascontiguousarray(a, dtype=t).tobytes())
    return h.hexdigest()
