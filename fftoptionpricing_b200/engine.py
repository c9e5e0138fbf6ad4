"""Host-side mirror of the reference's pricing / calibration interface over the C ABI.

``Engine`` owns one ``fop_handle_t`` (one GPU, one stream).  Its methods take numpy arrays (host
buffers, copied inside the call) or torch CUDA tensors / raw device pointers (no copies, work is
only enqueued) and have the same signatures as ``oracle.oracle.Oracle`` so that parity tests call
both sides identically.  The reference-named entry points (``CarrMadanCall``,
``FangOostCallPrice``, ``FSTSPrice``, ``getObjFn_arr`` ...) live in ``pricing.py`` /
``calibration.py`` on top of this class.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib

GAUSS, MERTON, CGMY, VG = 0, 1, 2, 3
TC_NONE, TC_CIR, TC_CIR_DIFFUSION = 0, 1, 2
CALL_PRICE, PUT_PRICE, CALL_DELTA, PUT_DELTA, CALL_GAMMA, PUT_GAMMA, CALL_THETA, PUT_THETA = range(8)
FSTS_PRICE, FSTS_DELTA, FSTS_GAMMA, FSTS_THETA = range(4)
PAYOFF_CALL, PAYOFF_PUT = 0, 1


class FopError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"fftop_b200 status {status}: {message}")
        self.status = status


def num_params(base: int, tc: int) -> int:
    return int(_lib.load_library().fop_model_num_params(_lib.fop_model_t(base, tc)))


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Buf:
    """Pointer + keep-alive for an argument that is a numpy array, torch tensor or None."""

    def __init__(self, x, dtype=np.float64, shape=None, writable=False):
        self.keep = None
        self.ptr = None
        self.is_device = False
        if x is None:
            return
        if _is_torch(x):
            import torch
            assert x.dtype == (torch.float64 if dtype == np.float64 else torch.complex128), x.dtype
            assert x.is_contiguous()
            self.keep = x
            self.ptr = C.c_void_p(x.data_ptr())
            self.is_device = x.is_cuda
            self.size = x.numel()
            return
        if writable:
            # an output buffer is never converted: a silent copy would be filled instead of the
            # caller's array
            if not isinstance(x, np.ndarray) or x.dtype != np.dtype(dtype):
                raise TypeError(f"output buffer must be a numpy array of dtype {np.dtype(dtype)}, got "
                                f"{type(x).__name__} {getattr(x, 'dtype', '')}")
            if not (x.flags.c_contiguous and x.flags.writeable):
                raise TypeError("output buffer must be C-contiguous and writeable")
            a = x
        else:
            a = np.ascontiguousarray(np.asarray(x, dtype=dtype))
        if shape is not None:
            a = a.reshape(shape)
        self.keep = a
        self.ptr = C.c_void_p(a.ctypes.data)
        self.size = a.size


class Engine:
    def __init__(self, device: int = 0):
        self._L = _lib.load_library()
        h = C.c_void_p()
        st = self._L.fop_create(C.byref(h), int(device))
        if st != 0:
            raise FopError(st, "fop_create failed: no usable CUDA device "
                               f"{device} (this library has no CPU fallback)")
        self._h = h
        self.device = device

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._L.fop_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != 0:
            raise FopError(st, self._L.fop_last_error(self._h).decode())

    @property
    def stream_ptr(self) -> int:
        return int(self._L.fop_get_stream(self._h) or 0)

    def synchronize(self):
        self._check(self._L.fop_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.fop_launch_count(self._h))

    @property
    def last_cos_contraction(self) -> str:
        """Contraction kernel of the last COS call: 'fp64-dmma', 'int8-tcgen05', 'fused' or 'none'."""
        return ("none", "fp64-dmma", "int8-tcgen05", "fused")[int(self._L.fop_last_cos_contraction(self._h))]

    def profile_enable(self, on: bool = True):
        """Bracket every kernel launch with CUDA events on the handle's stream."""
        self._check(self._L.fop_profile_enable(self._h, int(bool(on))))

    def profile_read(self):
        """{kernel name: (total ms, launches)} accumulated since the last read; synchronises."""
        names = C.create_string_buffer(4096)
        ms = (C.c_double * 64)()
        ln = (C.c_longlong * 64)()
        n = C.c_int()
        self._check(self._L.fop_profile_read(self._h, names, 4096, ms, ln, 64, C.byref(n)))
        parts = names.raw.split(b"\0")
        return {parts[i].decode(): (ms[i], int(ln[i])) for i in range(n.value)}

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._check(self._L.fop_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def microbench_fp64(self):
        v = (C.c_double * 4)()
        self._check(self._L.fop_microbench_fp64(self._h, v))
        return list(v)

    def microbench_int8(self):
        """Measured tcgen05.mma kind::i8 rate (POP/s) with shared-memory-resident operands."""
        v = C.c_double()
        self._check(self._L.fop_microbench_int8(self._h, C.byref(v)))
        return v.value

    @staticmethod
    def _params(base, tc, params):
        np_ = num_params(base, tc)
        if np_ == 0:
            raise FopError(1, f"unknown model ({base},{tc})")
        if _is_torch(params):
            assert params.numel() % np_ == 0
            return _Buf(params), params.numel() // np_
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64)).reshape(-1, np_)
        return _Buf(p), p.shape[0]

    @staticmethod
    def _out(out, shape):
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        b = _Buf(out, writable=True)
        assert b.size == int(np.prod(shape)), (b.size, shape)
        return out, b

    # ------------------------------------------------------------------ pricers
    def carr_madan(self, base, tc, params, N, ada, alpha, S0, r, T, is_put=False, out=None):
        pb, P = self._params(base, tc, params)
        out, ob = self._out(out, (P, N))
        self._check(self._L.fop_carr_madan(self._h, _lib.fop_model_t(base, tc), pb.ptr, P, int(N),
                                           float(ada), float(alpha), float(S0), float(r), float(T),
                                           int(bool(is_put)), ob.ptr))
        return out

    def cos(self, base, tc, params, K_desc, S0, r, T: Sequence[float], numU, kind=0, out=None):
        pb, P = self._params(base, tc, params)
        Kb = _Buf(K_desc)
        Tb = _Buf(T if _is_torch(T) else np.atleast_1d(np.asarray(T, dtype=np.float64)))
        out, ob = self._out(out, (P, Tb.size, Kb.size))
        self._check(self._L.fop_cos(self._h, _lib.fop_model_t(base, tc), pb.ptr, P, Kb.ptr, Kb.size,
                                    float(S0), float(r), Tb.ptr, Tb.size, int(numU), int(kind),
                                    ob.ptr))
        return out

    def cos_greeks(self, base, tc, params, K_desc, S0, r, T: Sequence[float], numU, kinds, out=None):
        """Several output kinds (price / delta / gamma / theta, call or put) from one pass over
        the characteristic function.  Returns [len(kinds)][P][M][J]."""
        pb, P = self._params(base, tc, params)
        Kb = _Buf(K_desc)
        Tb = _Buf(T if _is_torch(T) else np.atleast_1d(np.asarray(T, dtype=np.float64)))
        kk = (C.c_int * len(kinds))(*[int(k) for k in kinds])
        out, ob = self._out(out, (len(kinds), P, Tb.size, Kb.size))
        self._check(self._L.fop_cos_greeks(self._h, _lib.fop_model_t(base, tc), pb.ptr, P, Kb.ptr,
                                           Kb.size, float(S0), float(r), Tb.ptr, Tb.size, int(numU),
                                           kk, len(kinds), ob.ptr))
        return out

    def fsts(self, base, tc, params, N, xmax, strike, r, T, steps=1, american=False, payoff=0,
             kind=0, out=None):
        pb, P = self._params(base, tc, params)
        out, ob = self._out(out, (P, N))
        self._check(self._L.fop_fsts(self._h, _lib.fop_model_t(base, tc), pb.ptr, P, int(N),
                                     float(xmax), float(strike), float(r), float(T), int(steps),
                                     int(bool(american)), int(payoff), int(kind), ob.ptr))
        return out

    def objfn_cf(self, base, tc, params, u, phiHat, r, T, out=None):
        pb, P = self._params(base, tc, params)
        ub = _Buf(u)
        if _is_torch(phiHat):
            phb = _Buf(phiHat, dtype=np.complex128)
        else:
            ph = np.ascontiguousarray(np.asarray(phiHat, dtype=np.complex128)).view(np.float64)
            phb = _Buf(ph)
        out, ob = self._out(out, (P,))
        self._check(self._L.fop_objfn_cf(self._h, _lib.fop_model_t(base, tc), pb.ptr, P, ub.ptr,
                                         phb.ptr, ub.size, float(r), float(T), ob.ptr))
        return out

    # ------------------------------------------------------------------ multi-GPU loss gather
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL id (rank 0 creates it; hand it to the other ranks by any host means)."""
        buf = C.create_string_buffer(128)
        st = _lib.load_library().fop_comm_unique_id(buf)
        if st != 0:
            raise FopError(st, "fop_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        self._check(self._L.fop_comm_init(self._h, unique_id, int(rank), int(world)))
        self._world = int(world)

    def gather_losses(self, loss_local, out=None):
        """NCCL all-gather of this rank's per-set losses: returns [world * count], rank-major.
        numpy in -> numpy out; torch CUDA tensor in -> torch tensor out."""
        b = _Buf(loss_local)
        world = self._world
        if _is_torch(loss_local):
            import torch
            o = out if out is not None else torch.empty(world * b.size, dtype=torch.float64,
                                                        device=loss_local.device)
            ob = _Buf(o)
        else:
            o = out if out is not None else np.empty(world * b.size)
            ob = _Buf(o, writable=True)
        self._check(self._L.fop_gather_losses(self._h, b.ptr, b.size, ob.ptr))
        return o

    def gather_losses_async(self, loss_local, out):
        """Side-stream all-gather of device-resident losses (torch CUDA tensors), overlapped with
        whatever is enqueued next; ``comm_wait`` orders the handle's stream after it."""
        b, ob = _Buf(loss_local), _Buf(out)
        assert b.is_device and ob.is_device and ob.size == self._world * b.size
        self._check(self._L.fop_gather_losses_async(self._h, b.ptr, b.size, ob.ptr))
        return out

    def comm_wait(self):
        self._check(self._L.fop_comm_wait(self._h))

    def calibrate_cf(self, base, tc, lower, upper, u, phiHat, r, T, nests=25, iterations=1500,
                     tol=1e-12, seed=42, restarts=1):
        """On-device cuckoo search over the objective of ``objfn_cf`` (replaces the reference's
        cuckoo::optimize call, generateCharts.cpp:115).  Returns ``(params[restarts, np],
        fnval[restarts], iterations[restarts], best_restart)``."""
        npar = num_params(base, tc)
        lo = np.ascontiguousarray(np.asarray(lower, dtype=np.float64))
        hi = np.ascontiguousarray(np.asarray(upper, dtype=np.float64))
        assert lo.size == npar and hi.size == npar, (lo.size, hi.size, npar)
        ub = _Buf(u)
        ph = np.ascontiguousarray(np.asarray(phiHat, dtype=np.complex128))
        assert ph.size == ub.size
        prm = np.empty((restarts, npar))
        fv = np.empty(restarts)
        its = np.empty(restarts, dtype=np.int32)
        best = C.c_int(0)
        self._check(self._L.fop_calibrate_cf(
            self._h, _lib.fop_model_t(base, tc), lo.ctypes.data_as(_lib._dp), hi.ctypes.data_as(_lib._dp),
            ub.ptr, C.c_void_p(ph.ctypes.data), ub.size, float(r), float(T), int(nests),
            int(iterations), float(tol), int(seed), int(restarts), C.c_void_p(prm.ctypes.data),
            C.c_void_p(fv.ctypes.data), C.c_void_p(its.ctypes.data), C.byref(best)))
        return prm, fv, its, best.value

    def calibrate_cos(self, base, tc, lower, upper, K_desc, S0, r, T, numU, mkt, w=None, nests=25,
                      iterations=300, tol=0.0, seed=42, restarts=64):
        """Population cuckoo search over the PRICE-space objective of ``cos_loss`` (restarts x nests
        parameter sets priced per half-iteration by the batched COS kernels).  Returns ``(params[restarts,
        np], fnval[restarts], iterations, best_restart)``."""
        npar = num_params(base, tc)
        lo = np.ascontiguousarray(np.asarray(lower, dtype=np.float64))
        hi = np.ascontiguousarray(np.asarray(upper, dtype=np.float64))
        assert lo.size == npar and hi.size == npar, (lo.size, hi.size, npar)
        Kb = _Buf(K_desc)
        Tb = _Buf(np.atleast_1d(np.asarray(T, dtype=np.float64)))
        mb, wb = _Buf(mkt), _Buf(w)
        assert mb.size == Tb.size * Kb.size
        prm = np.empty((restarts, npar))
        fv = np.empty(restarts)
        its = np.empty(restarts, dtype=np.int32)
        best = C.c_int(0)
        self._check(self._L.fop_calibrate_cos(
            self._h, _lib.fop_model_t(base, tc), lo.ctypes.data_as(_lib._dp), hi.ctypes.data_as(_lib._dp),
            Kb.ptr, Kb.size, float(S0), float(r), Tb.ptr, Tb.size, int(numU), mb.ptr, wb.ptr, int(nests),
            int(iterations), float(tol), int(seed), int(restarts), C.c_void_p(prm.ctypes.data),
            C.c_void_p(fv.ctypes.data), C.c_void_p(its.ctypes.data), C.byref(best)))
        return prm, fv, int(its[0]), best.value

    def cos_loss(self, base, tc, params, K_desc, S0, r, T, numU, mkt, w=None, want_prices=False,
                 out=None, prices_out=None):
        pb, P = self._params(base, tc, params)
        Kb = _Buf(K_desc)
        Tb = _Buf(T if _is_torch(T) else np.atleast_1d(np.asarray(T, dtype=np.float64)))
        mb, wb = _Buf(mkt), _Buf(w)
        assert mb.size == Tb.size * Kb.size
        out, ob = self._out(out, (P,))
        pr, prb = (None, _Buf(None))
        if want_prices or prices_out is not None:
            pr, prb = self._out(prices_out, (P, Tb.size, Kb.size))
        self._check(self._L.fop_cos_loss(self._h, _lib.fop_model_t(base, tc), pb.ptr, P, Kb.ptr,
                                         Kb.size, float(S0), float(r), Tb.ptr, Tb.size, int(numU),
                                         mb.ptr, wb.ptr, ob.ptr, prb.ptr))
        return (out, pr) if want_prices else out

    def fft(self, x, inverse=False):
        """Unnormalised batched FFT over the last axis (fft.hpp:49-61).  numpy in -> numpy out
        (copy); a torch CUDA complex128 tensor is transformed in place."""
        if _is_torch(x):
            N = x.shape[-1]
            b = _Buf(x, dtype=np.complex128)
            self._check(self._L.fop_fft(self._h, b.ptr, x.numel() // N, N, int(bool(inverse))))
            return x
        z = np.array(x, dtype=np.complex128, order="C", copy=True)
        N = z.shape[-1]
        self._check(self._L.fop_fft(self._h, C.c_void_p(z.ctypes.data), z.size // N, N,
                                    int(bool(inverse))))
        return z

    def integrate_fft(self, fn_samples, outputMin, inputMin, inputMax, inverse=False, out=None):
        """integrateFFT / integrateIFFT (fft.hpp:96-110) of functions given by their samples on the
        uniform input grid, batched over the leading axes.  numpy in -> numpy out; a torch CUDA
        complex128 tensor in -> torch tensor out (``out`` may be the input)."""
        if _is_torch(fn_samples):
            import torch
            N = fn_samples.shape[-1]
            o = out if out is not None else torch.empty_like(fn_samples)
            ib, ob = _Buf(fn_samples, dtype=np.complex128), _Buf(o, dtype=np.complex128)
            self._check(self._L.fop_integrate_fft(self._h, ib.ptr, fn_samples.numel() // N, N,
                                                  float(outputMin), float(inputMin), float(inputMax),
                                                  int(bool(inverse)), ob.ptr))
            return o
        z = np.ascontiguousarray(np.asarray(fn_samples, dtype=np.complex128))
        N = z.shape[-1]
        o = np.empty_like(z)
        self._check(self._L.fop_integrate_fft(self._h, C.c_void_p(z.ctypes.data), z.size // N, N,
                                              float(outputMin), float(inputMin), float(inputMax),
                                              int(bool(inverse)), C.c_void_p(o.ctypes.data)))
        return o

    # ------------------------------------------------------------------ caller-sampled CF / payoff
    def cos_u_grid(self, K_desc, S0, numU):
        """u_k at which fop_cos_cf_samples expects cf(1j * u_k)."""
        K = np.ascontiguousarray(np.asarray(K_desc, dtype=np.float64))
        out = np.empty(int(numU))
        st = self._L.fop_cos_u_grid(K.ctypes.data_as(_lib._dp), K.size, float(S0), int(numU),
                                    out.ctypes.data_as(_lib._dp))
        if st != 0:
            raise FopError(st, "fop_cos_u_grid: bad arguments")
        return out

    def cos_cf_samples(self, cf_samples, K_desc, S0, r, T, numU, kind=0, out=None):
        """COS prices from sampled characteristic-function values cf_samples[P][M][numU] (complex):
        the reference's arbitrary-callable signature (FangOostGeneric, OptionPricing.h:325-333)."""
        Kb = _Buf(K_desc)
        Tb = _Buf(np.atleast_1d(np.asarray(T, dtype=np.float64)))
        if _is_torch(cf_samples):
            sb = _Buf(cf_samples, dtype=np.complex128)
            n = cf_samples.numel()
        else:
            z = np.ascontiguousarray(np.asarray(cf_samples, dtype=np.complex128))
            sb = _Buf(z.view(np.float64))
            n = z.size
        P = n // (Tb.size * int(numU))
        assert P * Tb.size * int(numU) == n, "cf_samples must be [P][M][numU]"
        out, ob = self._out(out, (P, Tb.size, Kb.size))
        self._check(self._L.fop_cos_cf_samples(self._h, sb.ptr, P, Kb.ptr, Kb.size, float(S0), float(r),
                                               Tb.ptr, Tb.size, int(numU), int(kind), ob.ptr))
        return out

    def carr_madan_u_grid(self, N, ada, alpha=1.5):
        out = np.empty(2 * int(N))
        self._L.fop_carr_madan_u_grid(int(N), float(ada), float(alpha), out.ctypes.data_as(_lib._dp))
        return out.view(np.complex128)

    def carr_madan_cf_samples(self, cf_samples, N, ada, alpha, S0, r, T, is_put=False, out=None):
        """cf_samples[P][N] (complex) = cf(alpha + 1 + 1j * j * ada): CarrMadanGeneric with an
        arbitrary callable (OptionPricing.h:582-607)."""
        z = np.ascontiguousarray(np.asarray(cf_samples, dtype=np.complex128)).reshape(-1, int(N))
        out, ob = self._out(out, (z.shape[0], int(N)))
        self._check(self._L.fop_carr_madan_cf_samples(self._h, C.c_void_p(z.ctypes.data), z.shape[0], int(N),
                                                      float(ada), float(alpha), float(S0), float(r), float(T),
                                                      int(bool(is_put)), ob.ptr))
        return out

    def fsts_u_grid(self, N, xmax):
        out = np.empty(int(N))
        self._L.fop_fsts_u_grid(int(N), float(xmax), out.ctypes.data_as(_lib._dp))
        return out

    def fsts_samples(self, cf_samples, payoff, N, xmax, r, T, steps=1, american=False, kind=0, asset=None,
                     out=None):
        """FSTSGeneric with arbitrary callables (OptionPricing.h:155-184): cf_samples[P][N] (complex) =
        cf(1j * w_k) for the step length T/steps, payoff[N] or [P][N], asset[N] for delta / gamma."""
        z = np.ascontiguousarray(np.asarray(cf_samples, dtype=np.complex128)).reshape(-1, int(N))
        P = z.shape[0]
        pay = np.ascontiguousarray(np.asarray(payoff, dtype=np.float64))
        per_set = 1 if pay.size == P * int(N) and P > 1 else 0
        assert pay.size in (int(N), P * int(N))
        ab = _Buf(asset)
        out, ob = self._out(out, (P, int(N)))
        self._check(self._L.fop_fsts_samples(self._h, C.c_void_p(z.ctypes.data), C.c_void_p(pay.ctypes.data), per_set,
                                             ab.ptr, P, int(N), float(xmax), float(r), float(T), int(steps),
                                             int(bool(american)), int(kind), ob.ptr))
        return out

    # ------------------------------------------------------------------ grids (host)
    def carr_madan_strikes(self, N, ada, S0):
        out = np.empty(N)
        self._L.fop_carr_madan_strikes(N, ada, S0, out.ctypes.data_as(_lib._dp))
        return out

    def fang_oost_strikes(self, xMin, xMax, S0, numX):
        out = np.empty(numX)
        self._L.fop_fang_oost_strikes(xMin, xMax, S0, numX, out.ctypes.data_as(_lib._dp))
        return out

    def strike_underlying(self, xMin, xMax, K, numX):
        out = np.empty(numX)
        self._L.fop_strike_underlying(xMin, xMax, K, numX, out.ctypes.data_as(_lib._dp))
        return out
