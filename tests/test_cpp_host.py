"""The C++ host layer (include/legume_b200.hpp: the compiled-language mirror of the reference's trait API over the
C ABI) builds here without a GPU and, on the GPU box, passes its own parity program (tests/cpp/test_host_api.cpp),
which restates the reference's tests for the path and checks the CUDA library against the CPU oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build():
    import oracle
    oracle.build()
    subprocess.run(["make", "-C", CPP], check=True, capture_output=True)
    return os.path.join(CPP, "test_host_api")


def test_cpp_host_layer_compiles_and_links():
    exe = _build()
    assert os.path.exists(exe)
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for s in ("lg_project", "lg_collapse_basic", "lg_knn_topk", "lg_pb_match", "lg_collect_matched_stat", "lg_optimize_batched"):
        assert s in syms, s


@pytest.mark.gpu
def test_cpp_host_api_parity_program():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:"), r.stdout
