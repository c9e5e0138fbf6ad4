"""Mirror of the reference's ``fft.h`` / ``fft.hpp``: ``fft``, ``ifft`` (fft.hpp:49-61, unnormalised)
and ``integrateFFT`` / ``integrateIFFT`` (fft.hpp:64-110).  The integrand is an ordinary Python
callable of the grid point, sampled on the host exactly as the reference does; the Simpson-weighted
transform runs on the device (fop_integrate_fft)."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import engine as E


def _engine(engine):
    from .pricing import default_engine
    return engine or default_engine()


def fft(x, engine: Optional[E.Engine] = None):
    """X_m = sum_j x_j e^{-2 pi i j m / N} (no 1/N), over the last axis."""
    return _engine(engine).fft(x, inverse=False)


def ifft(x, engine: Optional[E.Engine] = None):
    """x_j = sum_m X_m e^{+2 pi i j m / N} (no 1/N), over the last axis."""
    return _engine(engine).fft(x, inverse=True)


def _integrate(outputMin, inputMin, inputMax, fn: Callable, N, inverse, fnOutput, engine):
    inputDx = (inputMax - inputMin) / (N - 1)
    grid = inputMin + inputDx * np.arange(N)
    samples = np.array([fn(float(g)) for g in grid], dtype=np.complex128)
    out = _engine(engine).integrate_fft(samples, outputMin, inputMin, inputMax, inverse=inverse)
    if fnOutput is None:
        return out
    du = ((N - 1) * 2.0 * np.pi / (inputMax - inputMin)) / N
    return np.array([fnOutput(outputMin + i * du, v) for i, v in enumerate(out)])


def integrateFFT(xMin, uMin, uMax, fn: Callable, N, fnOutput: Optional[Callable] = None,
                 engine: Optional[E.Engine] = None):
    """fft.hpp:96-103: Fourier integral of fn over [uMin, uMax] on N grid points, outputs from xMin."""
    return _integrate(xMin, uMin, uMax, fn, N, False, fnOutput, engine)


def integrateIFFT(uMin, xMin, xMax, fn: Callable, N, fnOutput: Optional[Callable] = None,
                  engine: Optional[E.Engine] = None):
    """fft.hpp:104-110: the same with the inverse transform."""
    return _integrate(uMin, xMin, xMax, fn, N, True, fnOutput, engine)
