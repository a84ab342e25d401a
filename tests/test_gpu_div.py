"""div_by_rcp (the exact division-by-reciprocal used in the kernels' epilogues) == __fdiv_rn for every
float32 dividend, for the divisors in use.  Compiles a small test-only program with nvcc on the GPU box."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_div_by_rcp_is_ieee_division():
    nvcc = "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = os.path.join(tempfile.mkdtemp(), "div_check")
    cmd = [nvcc, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "gvcnn-tf_b200", "csrc"),
           os.path.join(ROOT, "tests", "cuda", "div_check.cu"), "-o", exe]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "mismatches 0 " in out.stdout, out.stdout + out.stderr
