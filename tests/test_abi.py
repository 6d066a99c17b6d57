"""The C-ABI library loads, exports every symbol include/kbgpu.h declares, the ctypes mirrors have the C
sizes, and — there being no GPU here — the product path fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from kube_batch_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(engine.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "kube_batch_b200", "csrc")])
    return engine.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "kbgpu.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z_]+)\s*\(", hdr))
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.kb_version()


def test_struct_sizes_match_c():
    src = r'''
#include <stdio.h>
#include "kbgpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(kb_snapshot), sizeof(kb_plugin_option), sizeof(kb_tier),
         sizeof(kb_plugin_conf), sizeof(kb_engine_opts), sizeof(kb_decision), sizeof(kb_stats), sizeof(kb_running));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [abi.kb_snapshot, abi.kb_plugin_option, abi.kb_tier, abi.kb_plugin_conf, abi.kb_engine_opts,
               abi.kb_decision, abi.kb_stats, abi.kb_running]
    assert sizes == [C.sizeof(m) for m in mirrors]
    import numpy as np
    assert np.dtype(abi.DECISION_DTYPE).itemsize == C.sizeof(abi.kb_decision) == 16


def test_status_strings(lib):
    assert lib.kb_status_str(0) == b"KB_OK"
    assert lib.kb_status_str(abi.KB_E_UNSUPPORTED_PLUGIN) == b"KB_E_UNSUPPORTED_PLUGIN"


def test_no_cpu_fallback_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.KbError) as ei:
        engine.Engine(device=0)
    assert ei.value.code == abi.KB_E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kube_batch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "kbo" not in re.findall(r"\bkbo\b", txt) and "oracle/" not in txt.replace("oracle/ ", ""), f
                assert "kb_oracle" not in txt, f
