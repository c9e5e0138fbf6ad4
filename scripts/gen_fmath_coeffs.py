#!/usr/bin/env python
"""Regenerates the exp polynomial of csrc/fmath.cuh (table fm_exp_q):

    q(r) ~ (e^r - 1 - r) / r^2   on |r| <= ln(2)/2,  degree 10,

as the Chebyshev interpolant of degree 10 (Chebyshev nodes of the interval, 60-digit arithmetic,
mpmath), converted to the monomial basis -- for a function this smooth the interpolant is within a
few percent of the minimax error (1.4e-18 here, far below one ulp of e^r = 1 + r + r^2 q(r)).
Prints the coefficients highest power first, the order the table stores them, and the maximum
difference from the table compiled into fmath.cuh.

    python scripts/gen_fmath_coeffs.py
The sin / cos / log / atan sets are Sun fdlibm's published minimax coefficients (k_sin.c, k_cos.c,
e_log.c, s_atan.c) and are not regenerated here."""
import os
import re

import mpmath as mp

mp.mp.dps = 60
DEG = 10
A = mp.log(2) / 2


def f(r):
    r = mp.mpf(r)
    if abs(r) < mp.mpf("1e-20"):
        return mp.mpf(1) / 2 + r / 6
    return (mp.exp(r) - 1 - r) / (r * r)


def chebyshev_monomial():
    n = DEG + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    vals = [f(A * t) for t in nodes]
    # Chebyshev coefficients c_j of the interpolant sum' c_j T_j(t)
    c = [2 / mp.mpf(n) * sum(vals[k] * mp.cos(mp.pi * j * (2 * k + 1) / (2 * n)) for k in range(n)) for j in range(n)]
    c[0] /= 2
    # T_j in the monomial basis (t), then t = r / A
    Tm = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        a, b = Tm[j - 1], Tm[j - 2]
        nxt = [mp.mpf(0)] + [2 * x for x in a]
        for i, x in enumerate(b):
            nxt[i] -= x
        Tm.append(nxt)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, x in enumerate(Tm[j]):
            mono[i] += c[j] * x
    return [mono[i] / A ** i for i in range(n)]   # coefficient of r^i


def table_in_source():
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fftoptionpricing_b200", "csrc",
                            "fmath.cuh")).read()
    m = re.search(r"FM_TABLE\(fm_exp_q,([^)]*)\)", src)
    return [float(x) for x in m.group(1).replace("\n", " ").split(",")]


if __name__ == "__main__":
    mono = chebyshev_monomial()
    hi_first = [float(x) for x in reversed(mono)]
    print("q(r), highest power first:")
    print(", ".join(repr(x) for x in hi_first))
    err = max(abs(sum(mono[i] * (A * t) ** i for i in range(DEG + 1)) - f(A * t))
              for t in [mp.mpf(k) / 500 - 1 for k in range(1001)])
    print("max |q - f| on the interval:", mp.nstr(err, 3))
    tab = table_in_source()
    rel = max(abs(a - b) / abs(b) for a, b in zip(hi_first, tab))
    print("max relative difference from fmath.cuh's table:", rel)
