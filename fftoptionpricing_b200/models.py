"""Characteristic-function model descriptors.

The reference takes an arbitrary C++ callable as characteristic function; device code honours the
closed set its tests and calibration driver use (include/fftop_b200.h, "model descriptor").  A
``CF`` carries the (base, time change) pair and one or more parameter sets.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .engine import CGMY, GAUSS, MERTON, TC_CIR, TC_CIR_DIFFUSION, TC_NONE, VG, num_params


@dataclass
class CF:
    base: int
    time_change: int
    params: np.ndarray  # [P][np]

    def __post_init__(self):
        n = num_params(self.base, self.time_change)
        self.params = np.ascontiguousarray(np.asarray(self.params, dtype=np.float64)).reshape(-1, n)

    @property
    def num_sets(self) -> int:
        return self.params.shape[0]


def black_scholes(sigma) -> CF:
    """exp(gaussLogCF(u, r - sigma^2/2, sigma) T)   (test.cpp:54, :570)"""
    return CF(GAUSS, TC_NONE, np.atleast_1d(sigma).reshape(-1, 1))


def merton(sigma, lam, muJ, sigJ) -> CF:
    """exp(mertonLogRNCF(u, lambda, muJ, sigJ, r, sigma) T)   (test.cpp:1119)"""
    return CF(MERTON, TC_NONE, np.column_stack(np.broadcast_arrays(*map(np.atleast_1d, (sigma, lam, muJ, sigJ)))))


def cgmy(sigma, C, G, M, Y) -> CF:
    """exp(cgmyLogRNCF(u, C, G, M, Y, r, sigma) T)   (test.cpp:731)"""
    return CF(CGMY, TC_NONE, np.column_stack(np.broadcast_arrays(*map(np.atleast_1d, (sigma, C, G, M, Y)))))


def variance_gamma(sigma, theta, nu) -> CF:
    """exp(vgLogRNCF(u, sigma, theta, nu, r) T): variance gamma (named by the north star; not in the
    reference -- textbook form, see include/fftop_b200.h)."""
    return CF(VG, TC_NONE, np.column_stack(np.broadcast_arrays(*map(np.atleast_1d, (sigma, theta, nu)))))


def heston(sigma, speed, adaV, rho, v0Hat) -> CF:
    """The reference's Heston: CIR-time-changed Brownian motion (test.cpp:1049-1058,
    generateCharts.cpp:231-250).  From the classical (b, a, c, rho, v0): sigma = sqrt(b),
    speed = a, adaV = c/sqrt(b), v0Hat = v0/b (test.cpp:1042-1048)."""
    return CF(GAUSS, TC_CIR, np.column_stack(np.broadcast_arrays(*map(np.atleast_1d, (sigma, speed, adaV, rho, v0Hat)))))


def time_changed(base_cf: CF, speed, adaV, rho, v0Hat, diffusion_only: bool = False) -> CF:
    """exp(r T u + cirLogMGF(-psi_0(u), speed, speed - adaV rho u sigma, adaV, T, v0Hat))
    (test.cpp:887-905)."""
    P = base_cf.num_sets
    extra = np.column_stack(np.broadcast_arrays(*map(np.atleast_1d, (speed, adaV, rho, v0Hat))))
    extra = np.broadcast_to(extra, (P, 4))
    return CF(base_cf.base, TC_CIR_DIFFUSION if diffusion_only else TC_CIR,
              np.hstack([base_cf.params, extra]))
