"""CPU test: the FFT index algebra of csrc/fft_smem.cuh (plans, digit reversal, DIF/DIT passes,
padding, twiddle power tree) emulated thread-by-thread on the host (csrc/host_fft_test.cu)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_passes_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "host_fft_test")
    src = os.path.join(ROOT, "fftoptionpricing_b200", "csrc", "host_fft_test.cu")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:]
