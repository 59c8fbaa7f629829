"""The C++ adapter header (include/omniswarm_b200_adapters.hpp) compiles with plain g++ against the C ABI; on a GPU the
adapter smoke program runs a database search, a cross-check match and a pose-graph solve through the adapters."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "omni-swarm_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "adapter_smoke")
    cmd = ["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cpp"),
           "-L", CSRC, "-lomniswarm_b200", f"-Wl,-rpath,{CSRC}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ missing")
def test_adapter_header_compiles_and_reports_no_device(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr      # without a GPU: create() returned OSB_ERR_NO_DEVICE


@pytest.mark.gpu
def test_adapters_on_gpu(gpu, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "adapters ok" in r.stdout, r.stdout + r.stderr
