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
