"""CPU test: accuracy of the branch-free fp64 elementary functions (csrc/fmath.cuh) against
long double libm -- exp/log <= 1 ulp, sincos <= 2 ulp for |x| up to 1e9, atan2 <= 2.5 ulp incl.
subnormal and huge operands (csrc/host_fmath_test.cu)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fmath_accuracy_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "host_fmath_test")
    src = os.path.join(ROOT, "fftoptionpricing_b200", "csrc", "host_fmath_test.cu")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:]
