"""GPU test: the C++ drop-in headers (include/fftop/*.h) compile with g++ and reproduce the
reference's own test cases (tests/cpp/test_dropin.cpp re-expresses test.cpp)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin(tmp_path):
    lib = os.path.join(ROOT, "fftoptionpricing_b200")
    exe = str(tmp_path / "test_dropin")
    r = subprocess.run(["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), "-o", exe,
                        "-L", lib, "-lfftop_b200", f"-Wl,-rpath,{lib}",
                        "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
