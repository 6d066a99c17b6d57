"""The C++ host mirror (kube_batch_b200/host/kbhost.hpp: api / cache / conf / framework.Session registration surface /
actions::allocate) driving libkbgpu.so through the C ABI, with the reference's own action test restated in C++
(tests/host/allocate_test.cpp <- pkg/scheduler/actions/allocate/allocate_test.go:38-212)."""
import os
import subprocess

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "host", "allocate_test")


def _build():
    if not os.path.exists(os.path.join(ROOT, "kube_batch_b200", "libkbgpu.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "kube_batch_b200", "csrc")])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s"])


def test_host_mirror_builds_and_fails_loudly_without_gpu():
    _build()
    if has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 3, (p.returncode, p.stderr)
    assert "no CPU fallback" in p.stderr and "KB_E_CUDA" in p.stderr


@pytest.mark.gpu
def test_reference_allocate_test_in_cpp():
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok   case") == 4


def test_cpp_flatten_matches_the_python_specification():
    """INTEGRATION.md "What Flatten must compute": the C++ mirror's Flatten (what the Go shim does) against
    kube_batch_b200/builder.py::SessionBuilder.flatten on the same objects — every numeric array, and the selector /
    toleration / host-port relations per (task, node).  Host code only: runs without a GPU."""
    import json

    import numpy as np

    from kube_batch_b200 import builder as B

    _build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host"), "-s", "flatten_dump"])
    got = json.loads(subprocess.run([os.path.join(ROOT, "tests", "host", "flatten_dump")], capture_output=True, text=True, check=True).stdout)

    G = 1e9
    b = B.SessionBuilder()
    b.add_node(B.Node("n0", {"cpu": 8, "memory": 32 * G, "pods": 10, "nvidia.com/gpu": 4}, labels={"zone": "a"}))
    b.add_node(B.Node("n1", {"cpu": 16, "memory": 64 * G, "pods": 20}, labels={"zone": "b"},
                      taints=[("dedicated", "batch", "NoSchedule"), ("soft", "x", "PreferNoSchedule")]))
    b.add_node(B.Node("n2", {"cpu": 4, "memory": 8 * G, "pods": 5, "nvidia.com/gpu": 2}, labels={"zone": "b"}, unschedulable=True,
                      memory_pressure=True))
    b.add_queue(B.Queue("q1", 1))
    b.add_queue(B.Queue("q2", 3))
    b.add_pod_group(B.PodGroup("ns", "old", "q1", 1))
    b.add_pod_group(B.PodGroup("ns", "pgA", "q1", 2))
    b.add_pod_group(B.PodGroup("ns", "pgB", "q2", 1, priority=7, creation=5))
    b.add_pod(B.Pod("ns", "r0", "n0", "Running", {"cpu": 2, "memory": 4 * G}, group="old", creation=1, host_ports=[("", "TCP", 8080)]))
    b.add_pod(B.Pod("ns", "r1", "n1", "Running", {"cpu": 3, "memory": 6 * G}, group="old", creation=2, deleting=True))
    b.add_pod(B.Pod("ns", "r2", "n2", "Running", {"cpu": 6, "memory": 1 * G}, group="old", creation=3))   # does not fit: stays off the node
    b.add_pod(B.Pod("ns", "a0", "", "Pending", {"cpu": 1, "memory": 1 * G}, group="pgA", creation=10, node_selector={"zone": "b"},
                    tolerations=[("dedicated", "Equal", "batch", "NoSchedule")]))
    b.add_pod(B.Pod("ns", "a1", "", "Pending", {"cpu": 1, "memory": 1 * G}, group="pgA", creation=11, node_selector={"zone": "b"},
                    host_ports=[("", "TCP", 8080)]))
    b.add_pod(B.Pod("ns", "b0", "", "Pending", {}, group="pgB", creation=12))
    b.add_pod(B.Pod("ns", "b1", "", "Pending", {"cpu": 0.5, "nvidia.com/gpu": 1}, group="pgB", creation=13, priority=5,
                    host_ports=[("10.0.0.1", "TCP", 8080), ("", "UDP", 53)], tolerations=[("", "Exists", "", "")]))
    s = b.flatten()

    assert (got["R"], got["N"], got["T"], got["J"], got["Q"]) == (s.R, s.N, s.T, s.J, s.Q)
    assert got["nodes"] == s.meta["nodes"] and got["tasks"] == s.meta["tasks"]
    for name in ("node_idle", "node_releasing", "node_used", "node_allocatable", "node_alloc_present", "node_alloc_cpu", "node_alloc_mem",
                 "node_nz_cpu", "node_nz_mem", "node_pods", "node_max_pods", "node_flags", "task_initreq", "task_resreq",
                 "task_res_present", "task_nz_cpu", "task_nz_mem", "task_flags", "task_prio", "task_ctime", "task_uid_rank", "job_task_off",
                 "job_min_avail", "job_ready0", "job_alloc0", "job_alloc0_present", "job_queue", "job_prio", "job_ctime", "queue_weight",
                 "queue_ctime"):
        want = np.asarray(getattr(s, name)).reshape(-1).astype(np.float64)
        have = np.asarray(got[name], dtype=np.float64)
        assert have.shape == want.shape and np.array_equal(have, want), (name, have.tolist(), want.tolist())
    T, N, W = s.T, s.N, s.W
    sel = np.ones((T, N), bool); tol = np.ones((T, N), bool); con = np.zeros((T, N), bool)
    for w in range(W):
        sel &= (s.task_sel_req[w][:, None] & ~s.node_labels[w][None, :]) == 0
        tol &= (s.node_taints[w][None, :] & ~s.task_tol[w][:, None]) == 0
        con |= (s.task_port_conflict[w][:, None] & s.node_ports[w][None, :]) != 0
    assert np.array_equal(np.asarray(got["rel_selector"], bool).reshape(T, N), sel)
    assert np.array_equal(np.asarray(got["rel_tolerated"], bool).reshape(T, N), tol)
    assert np.array_equal(np.asarray(got["rel_port_conflict"], bool).reshape(T, N), con)
    assert s.node_pods.tolist() == [1, 1, 0] and s.job_ready0.tolist() == [2, 0, 0]      # r2 counts for its job, not for n2
    # sanity of the fixture itself: a1's wildcard 8080/TCP collides with r0's on n0; b1's 10.0.0.1:8080 collides with that wildcard too
    assert con[1, 0] and con[3, 0] and not con[0].any()
