"""Regenerates the committed golden fixtures.

The reference (Go) cannot be executed in this image, so the fixtures are (a) the expected bind maps
transcribed from the reference's own action test
(/root/reference/pkg/scheduler/actions/allocate/allocate_test.go:51-144) and (b) oracle outputs for the
seeded BASELINE configs c1 / c2 — the oracle itself being pinned on the reference's known answers by
tests/test_oracle_golden.py — and (c) SHA-256 digests of the oracle's outcome for c2 / c3 / c4 (cycle_hashes.json: the
full tables would be megabytes), which bench.py checks on every rank at every GPU count and the -m gpu tests check at
full size.  Run from the repo root:  python tests/golden/make_golden.py [--no-c4]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kube_batch_b200 import synth  # noqa: E402
from oracle import kbo  # noqa: E402
from kube_batch_b200 import digest  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    json.dump({
        "source": "pkg/scheduler/actions/allocate/allocate_test.go:51-144",
        "case1": {"c1/p1": "n1", "c1/p2": "n1"},
        "case2": {"c2/p1": "n1", "c1/p1": "n1"},
    }, open(os.path.join(HERE, "allocate_test_expected.json"), "w"), indent=1)
    for name in ("c1", "c2"):
        snap, conf = synth.make(name)
        o = kbo.allocate(snap, conf)
        np.savez_compressed(os.path.join(HERE, f"{name}_oracle.npz"), decisions=o.decisions,
                            node_idle=o.node_idle, job_share=o.job_share, job_ready=o.job_ready)
        print(name, "tasks", snap.T, "allocated", o.result.tasks_allocated)
    hashes = {"source": "oracle/kb_oracle.cpp on synth.make(name); digests per kube_batch_b200/digest.py"}
    for name in ("c2", "c3") + (() if "--no-c4" in sys.argv else ("c4",)):
        snap, conf = synth.make(name)
        o = kbo.allocate(snap, conf, threads=os.cpu_count() or 1)
        hashes[name] = {"tasks": int(snap.T), "nodes": int(snap.N), "allocated": int(o.result.tasks_allocated),
                        "pipelined": int(o.result.tasks_pipelined), "visits": int(o.result.visits), "jobs_ready": int(o.result.jobs_ready),
                        "decisions": digest.decisions_digest(o.decisions),
                        "state": digest.state_digest(o.node_idle, o.node_releasing, o.job_ready, o.job_share)}
        print(name, hashes[name])
    if "--replicas" in sys.argv:          # the clusters ranks 1..7 schedule when bench.py runs N independent sessions (synth.make(name, replica))
        for r in range(1, 8):
            snap, conf = synth.make("c3", replica=r)
            o = kbo.allocate(snap, conf, threads=os.cpu_count() or 1)
            hashes[f"c3#{r}"] = {"tasks": int(snap.T), "nodes": int(snap.N), "allocated": int(o.result.tasks_allocated),
                                 "pipelined": int(o.result.tasks_pipelined), "visits": int(o.result.visits), "jobs_ready": int(o.result.jobs_ready),
                                 "decisions": digest.decisions_digest(o.decisions),
                                 "state": digest.state_digest(o.node_idle, o.node_releasing, o.job_ready, o.job_share)}
            print(f"c3#{r}", hashes[f"c3#{r}"])
    old = {}
    hp = os.path.join(HERE, "cycle_hashes.json")
    if os.path.exists(hp):
        old = json.load(open(hp))
    old.update(hashes)
    json.dump(old, open(hp, "w"), indent=1)
