"""TEST INFRASTRUCTURE ONLY: the reference's own formulas evaluated in
40-digit arithmetic with mpmath.

Used where the REFERENCE's double-precision result is itself rounding noise (e.g. the CIR clock at
adaV -> 0: (delta - kappa) speed / adaV^2 with delta = sqrt(kappa^2 + 2 psi adaV^2) cancels): there
`|gpu - ref| <= 1e-8 + 1e-10|ref|` cannot hold for ANY second implementation, and the meaningful
statement is "the GPU is as close to the exact value of the reference's formula as the reference
is".  Nothing here is used where the reference is well conditioned.
"""
from __future__ import annotations

import mpmath as mp
import numpy as np

mp.mp.dps = 40


def _gauss_log(u, mu, sig):
    return u * mu + sig * sig * u * u / 2


def levy_log_rn(base, p, u, rate, sig):
    """baseLogRNCF of oracle_models.h (gaussLogCF / mertonLogRNCF / cgmyLogRNCF)."""
    if base == 0:
        return _gauss_log(u, rate - sig * sig / 2, sig)
    if base == 1:
        lam, muJ, sJ = p[1], p[2], p[3]
        comp = lam * (mp.exp(muJ + sJ * sJ / 2) - 1)
        return _gauss_log(u, rate - sig * sig / 2 - comp, sig) + lam * (mp.exp(_gauss_log(u, muJ, sJ)) - 1)
    if base == 3:  # variance gamma [sigma, theta, nu], textbook (Madan-Carr-Chang 1998); no Brownian part
        s, th, nu = p[0], p[1], p[2]
        return u * (rate + mp.log(1 - th * nu - s * s * nu / 2) / nu) - mp.log(1 - u * th * nu - s * s * nu * u * u / 2) / nu
    C, G, M, Y = p[1], p[2], p[3], p[4]

    def k(z):
        if C == 0:
            return mp.mpf(0)
        return C * mp.gamma(-Y) * ((M - z) ** Y + (G + z) ** Y - M ** Y - G ** Y)
    return _gauss_log(u, rate - sig * sig / 2 - k(mp.mpf(1)), sig) + k(u)


def log_cf(base, tc, params, r, T, u):
    p = [mp.mpf(float(x)) for x in params]
    r, T, u = mp.mpf(float(r)), mp.mpf(float(T)), mp.mpc(u)
    sig = mp.mpf(0) if base == 3 else p[0]
    if tc == 0:
        return levy_log_rn(base, p, u, r, sig) * T
    nb = {0: 1, 1: 4, 2: 5, 3: 3}[base]
    speed, adaV, rho, v0 = p[nb:nb + 4]
    if tc == 1:
        psi = -levy_log_rn(base, p, u, mp.mpf(0), sig)
        extra = 0
    else:
        psi = -_gauss_log(u, -sig * sig / 2, sig)
        extra = levy_log_rn(base, p, u, mp.mpf(0), mp.mpf(0)) * T if base != 0 else 0
    kap = speed - adaV * rho * u * sig
    if adaV == 0:
        ek = 1 - mp.exp(-kap * T)
        b = psi * ek / kap
        c = (speed * psi / kap) * (T - ek / kap)
    else:
        delta = mp.sqrt(kap * kap + 2 * psi * adaV * adaV)
        e = mp.exp(-delta * T)
        b = 2 * psi * (1 - e) / (delta + kap + (delta - kap) * e)
        c = (speed / adaV ** 2) * (2 * mp.log(1 - (delta - kap) * (1 - e) / (2 * delta)) + (delta - kap) * T)
    return r * T * u - b * v0 - c + extra


def cos_call_prices(base, tc, params, K, S0, r, T, numU):
    """FangOostCallPrice (OptionPricing.h:362-382) for one parameter set, maturities T -> [M][J]."""
    K = [mp.mpf(float(k)) for k in K]
    x = [mp.log(mp.mpf(float(S0)) / k) for k in K]
    a, bb = x[0], x[-1]
    du, cp = mp.pi / (bb - a), 2 / (bb - a)
    J = len(K)
    uk = [du * k for k in range(numU)]
    vk = []
    for k, u in enumerate(uk):  # chiK / phiK with c = a, d = 0 (OptionPricing.h:284-300)
        chi = (mp.cos(u * (0 - a)) - mp.exp(a) + u * mp.sin(u * (0 - a))) / (1 + u * u)
        phi = (0 - a) if k == 0 else mp.sin(u * (0 - a)) / u
        vk.append(phi - chi)
    cosb = [[mp.cos(u * (xj - a)) for xj in x] for u in uk]
    sinb = [[mp.sin(u * (xj - a)) for xj in x] for u in uk]
    out = np.empty((len(T), J))
    for m, Tm in enumerate(T):
        disc = mp.exp(-mp.mpf(float(r)) * mp.mpf(float(Tm)))
        acc = [mp.mpf(0)] * J
        for k, u in enumerate(uk):
            D = mp.exp(log_cf(base, tc, params, r, Tm, mp.mpc(0, u))) * cp * vk[k]
            if k == 0:
                D = D / 2
            dr, di = D.real, D.imag
            ck, sk = cosb[k], sinb[k]
            acc = [acc[j] + dr * ck[j] - di * sk[j] for j in range(J)]
        for j in range(J):
            out[m, j] = float((acc[j] - 1) * disc * K[j] + mp.mpf(float(S0)))
    return out
