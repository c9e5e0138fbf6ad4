"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes loader for the CPU oracles.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``fftoptionpricing_b200``) never does.

Two flavours export the same C ABI (``oracle_abi.h``):

* ``reference``: ``oracle/_ref/libfftop_oracle_ref.so`` -- the reference's own headers compiled
  unchanged from ``/root/reference`` against ``oracle/shims`` (``ref_driver.cpp``).
* ``port``: ``oracle/libfftop_oracle_port.so`` -- independent restatement (``port/oracle_port.cpp``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libfftop_oracle_ref.so")
PORT_SO = os.path.join(_HERE, "libfftop_oracle_port.so")

GAUSS, MERTON, CGMY, VG = 0, 1, 2, 3
TC_NONE, TC_CIR, TC_CIR_DIFFUSION = 0, 1, 2

_dp = C.POINTER(C.c_double)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


def _arr(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        out = out.reshape(shape)
    return out


def num_params(base: int, tc: int) -> int:
    nb = {0: 1, 1: 4, 2: 5, 3: 3}.get(base, 0)
    if nb == 0 or tc not in (0, 1, 2):
        return 0
    return nb + (0 if tc == 0 else 4)


class Oracle:
    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.ora_flavor.restype = C.c_char_p
        L.ora_num_threads.restype = C.c_int
        i, d = C.c_int, C.c_double
        L.ora_carr_madan.argtypes = [i, i, _dp, i, i, d, d, d, d, d, i, _dp]
        L.ora_cos.argtypes = [i, i, _dp, i, _dp, i, d, d, _dp, i, i, i, _dp]
        L.ora_fsts.argtypes = [i, i, _dp, i, i, d, d, d, d, i, i, i, i, _dp]
        L.ora_objfn_cf.argtypes = [i, i, _dp, i, _dp, _dp, i, d, d, _dp]
        L.ora_cos_loss.argtypes = [i, i, _dp, i, _dp, i, d, d, _dp, i, i, _dp, _dp, _dp, _dp]
        L.ora_fft.argtypes = [_dp, i, i, i]
        L.ora_integrate_fft.argtypes = [_dp, i, i, d, d, d, i, _dp]
        L.ora_cf.argtypes = [i, i, _dp, d, d, d, d, _dp, _dp]
        L.ora_fo_estimate.argtypes = [_dp, _dp, i, d, d, d, d, d, i, _dp, i, _dp]
        L.ora_carr_madan_strikes.argtypes = [i, d, d, _dp]
        L.ora_fang_oost_strikes.argtypes = [d, d, d, i, _dp]
        L.ora_strike_underlying.argtypes = [d, d, d, i, _dp]

    @property
    def flavor(self) -> str:
        return self.lib.ora_flavor().decode()

    @property
    def num_threads(self) -> int:
        return int(self.lib.ora_num_threads())

    @staticmethod
    def _check(rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"oracle {what} failed with status {rc}")

    def _params(self, base, tc, params):
        np_ = num_params(base, tc)
        p = _arr(params).reshape(-1, np_)
        return p, p.shape[0]

    def carr_madan(self, base, tc, params, N, ada, alpha, S0, r, T, is_put=False):
        p, P = self._params(base, tc, params)
        out = np.empty((P, N))
        self._check(self.lib.ora_carr_madan(base, tc, _ptr(p), P, N, ada, alpha, S0, r, T,
                                            int(is_put), _ptr(out)), "carr_madan")
        return out

    def cos(self, base, tc, params, K_desc, S0, r, T: Sequence[float], numU, kind=0):
        p, P = self._params(base, tc, params)
        K = _arr(K_desc)
        Tm = _arr(np.atleast_1d(T))
        out = np.empty((P, Tm.size, K.size))
        self._check(self.lib.ora_cos(base, tc, _ptr(p), P, _ptr(K), K.size, S0, r, _ptr(Tm),
                                     Tm.size, numU, kind, _ptr(out)), "cos")
        return out

    def fsts(self, base, tc, params, N, xmax, strike, r, T, steps=1, american=False,
             payoff=0, kind=0):
        p, P = self._params(base, tc, params)
        out = np.empty((P, N))
        self._check(self.lib.ora_fsts(base, tc, _ptr(p), P, N, xmax, strike, r, T, steps,
                                      int(american), payoff, kind, _ptr(out)), "fsts")
        return out

    def objfn_cf(self, base, tc, params, u, phiHat, r, T):
        p, P = self._params(base, tc, params)
        uu = _arr(u)
        ph = np.ascontiguousarray(np.asarray(phiHat, dtype=np.complex128)).view(np.float64)
        loss = np.empty(P)
        self._check(self.lib.ora_objfn_cf(base, tc, _ptr(p), P, _ptr(uu), _ptr(ph), uu.size, r,
                                          T, _ptr(loss)), "objfn_cf")
        return loss

    def cos_loss(self, base, tc, params, K_desc, S0, r, T, numU, mkt, w=None,
                 want_prices=False):
        p, P = self._params(base, tc, params)
        K = _arr(K_desc)
        Tm = _arr(np.atleast_1d(T))
        mk = _arr(mkt, (Tm.size, K.size))
        ww = None if w is None else _arr(w, (Tm.size, K.size))
        loss = np.empty(P)
        prices = np.empty((P, Tm.size, K.size)) if want_prices else None
        self._check(self.lib.ora_cos_loss(base, tc, _ptr(p), P, _ptr(K), K.size, S0, r, _ptr(Tm),
                                          Tm.size, numU, _ptr(mk), _ptr(ww), _ptr(loss),
                                          _ptr(prices)), "cos_loss")
        return (loss, prices) if want_prices else loss

    def fft(self, x, inverse=False):
        z = np.ascontiguousarray(np.asarray(x, dtype=np.complex128))
        shape = z.shape
        z2 = z.reshape(-1, shape[-1]).copy()
        flat = z2.view(np.float64)
        self._check(self.lib.ora_fft(_ptr(flat), z2.shape[0], shape[-1], int(inverse)), "fft")
        return z2.reshape(shape)

    def integrate_fft(self, fn_samples, outputMin, inputMin, inputMax, inverse=False):
        z = np.ascontiguousarray(np.asarray(fn_samples, dtype=np.complex128))
        shape = z.shape
        z2 = z.reshape(-1, shape[-1])
        out = np.empty_like(z2)
        self._check(self.lib.ora_integrate_fft(_ptr(z2.view(np.float64)), z2.shape[0], shape[-1],
                                               float(outputMin), float(inputMin), float(inputMax),
                                               int(inverse), _ptr(out.view(np.float64))),
                    "integrate_fft")
        return out.reshape(shape)

    def cf(self, base, tc, params, r, T, u: complex):
        p = _arr(params)
        cf = np.empty(2)
        lcf = np.empty(2)
        self._check(self.lib.ora_cf(base, tc, _ptr(p), r, T, float(np.real(u)),
                                    float(np.imag(u)), _ptr(cf), _ptr(lcf)), "cf")
        return complex(cf[0], cf[1]), complex(lcf[0], lcf[1])

    def fo_estimate(self, strikes, options, stock, r, t, minStrike, maxStrike, N, u):
        k, o, uu = _arr(strikes), _arr(options), _arr(u)
        out = np.empty(2 * uu.size)
        self._check(self.lib.ora_fo_estimate(_ptr(k), _ptr(o), k.size, stock, r, t, minStrike,
                                             maxStrike, N, _ptr(uu), uu.size, _ptr(out)), "fo_estimate")
        return out.view(np.complex128).copy()

    def carr_madan_strikes(self, N, ada, S0):
        out = np.empty(N)
        self.lib.ora_carr_madan_strikes(N, ada, S0, _ptr(out))
        return out

    def fang_oost_strikes(self, xMin, xMax, S0, numX):
        out = np.empty(numX)
        self.lib.ora_fang_oost_strikes(xMin, xMax, S0, numX, _ptr(out))
        return out

    def strike_underlying(self, xMin, xMax, K, numX):
        out = np.empty(numX)
        self.lib.ora_strike_underlying(xMin, xMax, K, numX, _ptr(out))
        return out


def load(flavor: str = "best") -> Oracle:
    """flavor: 'reference', 'port' or 'best' (reference if built, else port)."""
    if flavor in ("reference", "best") and os.path.exists(REF_SO):
        return Oracle(REF_SO)
    if flavor == "reference":
        raise FileNotFoundError(f"{REF_SO} not built (needs /root/reference; run oracle/build.py)")
    if os.path.exists(PORT_SO):
        return Oracle(PORT_SO)
    raise FileNotFoundError(f"{PORT_SO} not built; run `python oracle/build.py`")


def available(flavor: str) -> bool:
    return os.path.exists(REF_SO if flavor == "reference" else PORT_SO)
