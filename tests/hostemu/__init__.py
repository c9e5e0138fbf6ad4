"""TEST INFRASTRUCTURE ONLY: builds and loads the CPU emulation of the device CF code (emu.cu)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libfop_hostemu.so")
_dp = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(HERE, "emu.cu")
    inc = os.path.join(ROOT, "fftoptionpricing_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("fmath.cuh", "cplx.cuh", "cf_models.cuh", "cos_coeff.cuh",
                                                   "cos_host_tables.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return None
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                        "-fPIC,-fopenmp,-mfma", "-shared", "-I", inc, "-o", SO, src],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostemu build failed:\n" + r.stderr[-3000:])
    return SO


class HostEmu:
    def __init__(self):
        so = build()
        if so is None:
            raise FileNotFoundError("nvcc not available")
        self.lib = C.CDLL(so)
        i, d = C.c_int, C.c_double
        self.lib.emu_cos.argtypes = [i, i, _dp, i, _dp, i, d, d, _dp, i, i, i, i, i, _dp]
        self.lib.emu_logcf.argtypes = [i, i, _dp, d, d, d, d, _dp]
        self.lib.emu_q8_store.argtypes = [_dp, _dp, i, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def q8_store(self, re, im):
        """Digit planes of the int8 contraction for n complex values: (q_re, q_im, rowflag, planes)."""
        re = np.ascontiguousarray(re, dtype=np.float64)
        im = np.ascontiguousarray(im, dtype=np.float64)
        n = re.size
        planes = np.zeros((n, 6, 2), dtype=np.uint8)
        flag = np.zeros(n, dtype=np.uint8)
        qr = np.zeros(n, dtype=np.int64)
        qi = np.zeros(n, dtype=np.int64)
        rc = self.lib.emu_q8_store(re.ctypes.data_as(_dp), im.ctypes.data_as(_dp), n, planes.ctypes.data,
                                   flag.ctypes.data, qr.ctypes.data, qi.ctypes.data)
        assert rc == 0
        return qr, qi, flag, planes

    def cos(self, base, tc, params, K, S0, r, T, numU, kind=0, use_tgrid=True, fast=True):
        K = np.ascontiguousarray(K, dtype=np.float64)
        T = np.ascontiguousarray(np.atleast_1d(T), dtype=np.float64)
        p = np.ascontiguousarray(params, dtype=np.float64)
        nb = {0: 1, 1: 4, 2: 5, 3: 3}[base]
        P = p.size // nb if tc == 0 else p.size // (nb + 4)
        out = np.empty((P, T.size, K.size))
        rc = self.lib.emu_cos(base, tc, p.ctypes.data_as(_dp), P, K.ctypes.data_as(_dp), K.size, S0, r,
                              T.ctypes.data_as(_dp), T.size, numU, kind, int(use_tgrid), int(fast),
                              out.ctypes.data_as(_dp))
        assert rc == 0
        return out

    def logcf(self, base, tc, params, r, T, u: complex):
        p = np.ascontiguousarray(params, dtype=np.float64)
        o = np.empty(2)
        assert self.lib.emu_logcf(base, tc, p.ctypes.data_as(_dp), r, T, u.real, u.imag,
                                  o.ctypes.data_as(_dp)) == 0
        return complex(o[0], o[1])
