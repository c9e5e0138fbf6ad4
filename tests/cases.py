"""Case catalogue shared by oracle/gen_golden.py and the parity tests.

Every case is a dict ``{"name", "fn", "kw", "ref"}``: ``fn`` names an oracle / product entry
point, ``kw`` its keyword arguments (JSON-serialisable), ``ref`` the reference test it re-expresses
(``/root/reference/test.cpp`` line) and, where the reference asserts one, ``kat``: a list of
``[flat_index, expected, catch_epsilon]`` known answers.

Model parameter layout: see include/fftop_b200.h (base params then [speed, adaV, rho, v0Hat]).
"""
from __future__ import annotations

import math

import numpy as np

GAUSS, MERTON, CGMY, VG = 0, 1, 2, 3
TC_NONE, TC_CIR, TC_CIR_DIFFUSION = 0, 1, 2
SEED = 20240611


# ------------------------------------------------------------------ grids (OptionPricing.h:84-138)
def fang_oost_strikes(xMin, xMax, S0, numX):
    dx = (xMax - xMin) / (numX - 1.0)
    return S0 * np.exp(-(xMin + dx * np.arange(numX)))


def carr_madan_strikes(N, ada, S0):
    b = math.pi / ada
    lam = 2.0 * b / N
    return S0 * np.exp(-b + lam * np.arange(N))


def strike_underlying(xMin, xMax, K, numX):
    dx = (xMax - xMin) / (numX - 1.0)
    return K * np.exp(xMin + dx * np.arange(numX))


# ------------------------------------------------------------------ closed-form BS (test.cpp:13-45)
def _ncdf(x):
    return 0.5 + 0.5 * math.erf(x / math.sqrt(2.0))


def bs_call(S0, discount, k, sigma, T):
    ss = math.sqrt(T) * sigma
    d1 = math.log(S0 / (discount * k)) / ss + ss * .5
    return S0 * _ncdf(d1) - k * discount * _ncdf(d1 - ss)


def bs_put(S0, discount, k, sigma, T):
    return bs_call(S0, discount, k, sigma, T) - S0 + k * discount


# ------------------------------------------------------------------ Heston set #0 (test.cpp:1035-1048)
def heston0():
    b, a, c, rho, v0 = .0398, 1.5768, .5751, -.5711, .0175
    return [math.sqrt(b), a, c / math.sqrt(b), rho, v0 / b]


# ------------------------------------------------------------------ synthetic batches
def heston_batch(P, seed=SEED):
    """(sig, speed, adaV, rho, v0Hat) ~ U over the shrunk generateCharts.cpp:217-223 boxes; set 0 =
    test.cpp:1035-1048."""
    rng = np.random.default_rng(seed)
    lo = np.array([0.05, 0.1, 0.1, -0.95, 0.05])
    hi = np.array([0.6, 2.0, 4.0, 0.95, 2.0])
    p = lo + (hi - lo) * rng.random((P, 5))
    p[0] = heston0()
    return p


def merton_batch(P, seed=SEED):
    """(sigma, lambda, muJ, sigJ) ~ U([.05,.6]x[0,2]x[-1,.5]x[.05,1.5]); set 0 = d'Halluin et al."""
    rng = np.random.default_rng(seed + 1)
    lo = np.array([0.05, 0.0, -1.0, 0.05])
    hi = np.array([0.6, 2.0, 0.5, 1.5])
    p = lo + (hi - lo) * rng.random((P, 4))
    p[0] = [0.15, 0.1, -0.9, 0.45]
    return p


def cgmy_batch(P, seed=SEED):
    """(sigma, C, G, M, Y) ~ U([0,.6]x[.01,2]x[1,15]x[2.6,15]x[.1,1.9]\\{1}); set 0 = test.cpp:835-842."""
    rng = np.random.default_rng(seed + 2)
    lo = np.array([0.0, 0.01, 1.0, 2.6, 0.1])
    hi = np.array([0.6, 2.0, 15.0, 15.0, 1.9])
    p = lo + (hi - lo) * rng.random((P, 5))
    y = p[:, 4]
    y[np.abs(y - 1.0) < 0.02] = 1.05
    p[0] = [0.2, 1.0, 1.4, 2.5, 1.4]
    return p


def vg_batch(P, seed=SEED):
    """(sigma, theta, nu) ~ U([.05,.5]x[-.4,.2]x[.05,.8]); set 0 = Madan-Carr-Chang (.12, -.14, .2)."""
    rng = np.random.default_rng(seed + 3)
    lo = np.array([0.05, -0.4, 0.05])
    hi = np.array([0.5, 0.2, 0.8])
    p = lo + (hi - lo) * rng.random((P, 3))
    p[0] = [0.12, -0.14, 0.2]
    return p


C5_T = [1.0 / 12, .25, .5, .75, 1.0, 1.5, 2.0, 3.0]


def c5_strikes(S0=100.0):
    """64 strikes S0*linspace(1.5,0.5,64) descending plus the synthetic end points S0*50, S0*0.03
    (convention of generateCharts.cpp:43-48) -> 66 evaluated, 64 counted."""
    k = S0 * np.linspace(1.5, 0.5, 64)
    return np.concatenate([[S0 * 50.0], k, [S0 * 0.03]])


def c5_weights():
    w = np.ones((len(C5_T), 66))
    w[:, 0] = 0.0
    w[:, -1] = 0.0
    return w


# ------------------------------------------------------------------ the catalogue
def catalogue():
    cases = []
    add = cases.append
    r, sig, T, S = .05, .3, 1.0, 50.0
    lo, hi = int(1024 * .3), int(1024 * .7)

    # --- FSTS vs Black-Scholes (test.cpp:47-338)
    for nm, payoff, kind, Tm, xmax, line in [
        ("FSTSCall", 0, 0, 1.0, 5.0, 47), ("FSTSCallDelta", 0, 1, 1.0, 5.0, 82),
        ("FSTSCallTheta", 0, 3, 1.0, 5.0, 117), ("FSTSCallGamma", 0, 2, 1.0, 5.0, 155),
        ("FSTSPut", 1, 0, 1.0, 5.0, 193), ("FSTSPutDelta", 1, 1, 1.0, 5.0, 230),
        ("FSTSPutTheta", 1, 3, 1.0, 5.0, 266), ("FSTSPutLowT", 1, 0, .25, 5.0 * math.sqrt(.25), 304),
    ]:
        add(dict(name=nm, fn="fsts", ref=f"test.cpp:{line}",
                 kw=dict(base=GAUSS, tc=0, params=[sig], N=1024, xmax=xmax, strike=S, r=r, T=Tm,
                         steps=1, american=False, payoff=payoff, kind=kind)))
    # --- COS vs Black-Scholes (test.cpp:342-531, 591-617)
    for nm, kind, line in [("FangOosterleeCall", 0, 342), ("FangOosterleeCallDelta", 2, 370),
                           ("FangOosterleeCallTheta", 6, 398), ("FangOosterleeCallGamma", 4, 425),
                           ("FangOosterleePutGamma", 5, 452), ("FangOosterleePutTheta", 7, 477),
                           ("FangOosterleePutDelta", 3, 504), ("FangOosterleePut", 1, 591)]:
        add(dict(name=nm, fn="cos", ref=f"test.cpp:{line}",
                 kw=dict(base=GAUSS, tc=0, params=[sig], K=("fang_oost", -5.0, 5.0, S, 1024),
                         S0=S, r=r, T=[1.0], numU=64, kind=kind)))
    add(dict(name="FangOosterleeCallFewK", fn="cos", ref="test.cpp:534",
             kw=dict(base=GAUSS, tc=0, params=[sig], K=[7500, 60, 50, 40, .3], S0=S, r=r, T=[1.0],
                     numU=64, kind=0)))
    add(dict(name="FangOosterleeCallFewKLowT", fn="cos", ref="test.cpp:562",
             kw=dict(base=GAUSS, tc=0, params=[sig], K=[7500, 60, 50, 40, .3], S0=S, r=r,
                     T=[.25], numU=256, kind=0)))
    # --- Carr-Madan vs Black-Scholes (test.cpp:619-708)
    for nm, Tm, put, line in [("CarrMadanCall", 1.0, False, 619), ("CarrMadanCallLowT", .25, False, 649),
                              ("CarrMadanPut", 1.0, True, 680)]:
        add(dict(name=nm, fn="carr_madan", ref=f"test.cpp:{line}",
                 kw=dict(base=GAUSS, tc=0, params=[sig], N=1024, ada=.25, alpha=1.5, S0=S, r=r,
                         T=Tm, is_put=put)))
    # --- CGMY (test.cpp:712-865)
    add(dict(name="FangOosterleeCallCGMY", fn="cos", ref="test.cpp:712", kat=[[1, 16.212478, 1e-4]],
             kw=dict(base=CGMY, tc=0, params=[0.0, 16.97, 7.08, 29.97, .6442], K=[7500, 98, .3],
                     S0=90.0, r=.06, T=[.25], numU=64, kind=0)))
    add(dict(name="FangOosterleeCallCGMYWitht=1", fn="cos", ref="test.cpp:741",
             kat=[[1, 19.812948843, 1e-4]],
             kw=dict(base=CGMY, tc=0, params=[0.0, 1.0, 5.0, 5.0, .5], K=[7500, 100.0, .3], S0=100.0,
                     r=.1, T=[1.0], numU=64, kind=0)))
    add(dict(name="FangOosterleeCallCGMYReducesToBS", fn="cos", ref="test.cpp:770",
             kat=[[1, bs_call(500.0, math.exp(-.4 * .25), 500.0, .2, .25), 1e-4]],
             kw=dict(base=CGMY, tc=0, params=[.2, 0.0, 1.4, 2.5, 1.4], K=[7500, 500.0, .3], S0=500.0,
                     r=.4, T=[.25], numU=256, kind=0)))
    add(dict(name="FangOosterleePutCGMYWithVol", fn="cos", ref="test.cpp:799",
             kat=[[1, 108.445317144, 1e-10]],
             kw=dict(base=CGMY, tc=0, params=[.2, 1.0, 1.4, 2.5, 1.4], K=[7500, 500.0, .3], S0=500.0,
                     r=.4, T=[.25], numU=256, kind=1)))
    add(dict(name="CarrMadanCGMY", fn="carr_madan", ref="test.cpp:831",
             kat=[[512, 108.4614527579, 1e-10]],
             kw=dict(base=CGMY, tc=0, params=[.2, 1.0, 1.4, 2.5, 1.4], N=1024, ada=.25, alpha=1.5,
                     S0=500.0, r=.4, T=.25, is_put=True)))
    # --- CIR-time-changed CGMY (test.cpp:872-966)
    cg = [.2, 1.0, 1.4, 2.5, .6]
    for nm, tc, extra, kat in [
        ("CarrMadanCGMYPut.corr", TC_CIR, [.5, .2, -1.0, 1.05], [[512, 74.9251, 1e-4]]),
        ("CarrMadanCGMYPut.stoch", TC_CIR, [.5, .2, 0.0, 1.05], [[512, 75.0101, 1e-4]]),
        ("CarrMadanCGMYPut.base", TC_NONE, [], [[512, 72.9703833045, 1e-10]]),
        ("CarrMadanCGMYPut.corrBM", TC_CIR_DIFFUSION, [.5, .2, -1.0, 1.05], []),
    ]:
        add(dict(name=nm, fn="carr_madan", ref="test.cpp:872", kat=kat,
                 kw=dict(base=CGMY, tc=tc, params=cg + extra, N=1024, ada=.25, alpha=1.5, S0=500.0,
                         r=.04, T=.25, is_put=True)))
    # --- Heston (test.cpp:970-1107, 1132-1180)
    h = heston0()
    add(dict(name="CarrMadanHeston.cgmy0", fn="carr_madan", ref="test.cpp:970",
             kat=[[512, 5.78515545, 1e-4]],
             kw=dict(base=CGMY, tc=TC_CIR, params=[h[0], 0.0, 1.0, 1.0, .5] + h[1:], N=1024, ada=.25,
                     alpha=1.5, S0=100.0, r=0.0, T=1.0, is_put=True)))
    add(dict(name="CarrMadanHeston.bm", fn="carr_madan", ref="test.cpp:998",
             kat=[[512, 5.78515545, 1e-4]],
             kw=dict(base=GAUSS, tc=TC_CIR, params=h, N=1024, ada=.25, alpha=1.5, S0=100.0, r=0.0,
                     T=1.0, is_put=True)))
    add(dict(name="FangOostHeston", fn="cos", ref="test.cpp:1033", kat=[[1, 5.78515545, 1e-4]],
             kw=dict(base=GAUSS, tc=TC_CIR, params=h, K=[5000, 100, .3], S0=100.0, r=0.0, T=[1.0],
                     numU=256, kind=0)))
    add(dict(name="FangOostHestonSubsetMerton", fn="cos", ref="test.cpp:1069",
             kat=[[1, 5.78515545, 1e-4]],
             kw=dict(base=MERTON, tc=TC_CIR, params=[h[0], 0.0, 1.0, 1.0] + h[1:], K=[5000, 100, .3],
                     S0=100.0, r=0.0, T=[1.0], numU=256, kind=0)))
    sl = math.sqrt(.05)
    add(dict(name="FangOostMerton", fn="cos", ref="test.cpp:1109", kat=[[1, 5.9713, 1e-4]],
             kw=dict(base=MERTON, tc=0, params=[sl, 1.0, -sl * sl * .5, sl], K=[5000, 35, .3],
                     S0=38.0, r=.1, T=[.5], numU=256, kind=0)))
    add(dict(name="FSTSHeston", fn="fsts", ref="test.cpp:1132", kat=[[512, 6.09466, 1e-3]],
             kw=dict(base=GAUSS, tc=TC_CIR, params=h, N=1024, xmax=5.0, strike=100.0, r=0.0, T=1.0,
                     steps=1, american=False, payoff=0, kind=0)))
    add(dict(name="FangOostDegenerateBS", fn="cos", ref="test.cpp:1184",
             kat=[[1, bs_call(100.0, math.exp(-.03 * .25), 100.0, .3, .25), 1e-4]],
             kw=dict(base=GAUSS, tc=TC_CIR, params=[.3, .4, 0.0, .3, 1.0], K=[5000, 100, .3],
                     S0=100.0, r=.03, T=[.25], numU=256, kind=0)))
    # --- calibration objective (OptionCalibration.h:185-201; generateCharts.cpp:149,231-250)
    uarr = [float(x) for x in np.linspace(-20.0, 20.0, 15)]
    for nm, base, tc, prm, Tm in [
        ("objfn.heston", GAUSS, TC_CIR, heston_batch(6).tolist(), 1.0),
        ("objfn.cgmy", CGMY, 0, cgmy_batch(6).tolist(), 1.0),
        ("objfn.merton", MERTON, 0, merton_batch(6).tolist(), .25),
        ("objfn.mertoncorr", MERTON, TC_CIR,
         np.hstack([merton_batch(6), heston_batch(6)[:, 1:]]).tolist(), .25),
    ]:
        add(dict(name=nm, fn="objfn_cf", ref="OptionCalibration.h:186",
                 kw=dict(base=base, tc=tc, params=prm, u=uarr, phiHat_from_set=0, r=0.0, T=Tm)))
    # --- BASELINE.json configs at oracle-friendly sizes
    add(dict(name="C2.heston_cos_1024", fn="cos", ref="BASELINE.json configs[1]",
             kw=dict(base=GAUSS, tc=TC_CIR, params=heston_batch(5).tolist(),
                     K=("fang_oost", -5.0, 5.0, 100.0, 1024), S0=100.0, r=0.0, T=[1.0], numU=256,
                     kind=0)))
    add(dict(name="C3.merton_american_put", fn="fsts", ref="BASELINE.json configs[2]",
             kw=dict(base=MERTON, tc=0, params=merton_batch(3).tolist(), N=4096, xmax=7.5,
                     strike=100.0, r=.05, T=.25, steps=256, american=True, payoff=1, kind=0)))
    add(dict(name="C3.merton_european_put_multistep", fn="fsts", ref="SURVEY.md A.4",
             kw=dict(base=MERTON, tc=0, params=merton_batch(2).tolist(), N=1024, xmax=7.5,
                     strike=100.0, r=.05, T=.25, steps=16, american=False, payoff=1, kind=0)))
    add(dict(name="C4.cgmy_carr_madan_65536", fn="carr_madan", ref="BASELINE.json configs[3]",
             kw=dict(base=CGMY, tc=0, params=cgmy_batch(3).tolist(), N=65536, ada=.25, alpha=1.5,
                     S0=500.0, r=.4, T=.25, is_put=False)))
    add(dict(name="C5.heston_sweep", fn="cos_loss", ref="BASELINE.json configs[4]",
             kw=dict(base=GAUSS, tc=TC_CIR, params=heston_batch(6).tolist(), K=("c5",), S0=100.0,
                     r=0.0, T=C5_T, numU=256, mkt_from_set=0, w=("c5",))))
    add(dict(name="C5.heston_cos_prices", fn="cos", ref="BASELINE.json configs[4]",
             kw=dict(base=GAUSS, tc=TC_CIR, params=heston_batch(6).tolist(), K=("c5",), S0=100.0,
                     r=0.0, T=C5_T, numU=256, kind=0)))
    return cases


def resolve_K(spec):
    if isinstance(spec, (list, tuple)) and len(spec) and spec[0] == "fang_oost":
        return fang_oost_strikes(*spec[1:])
    if isinstance(spec, (list, tuple)) and len(spec) and spec[0] == "c5":
        return c5_strikes()
    return np.asarray(spec, dtype=np.float64)


def run_case(impl, case):
    """Run one catalogue case on `impl` (an oracle.Oracle or the product's python binding --
    both expose carr_madan / cos / fsts / objfn_cf / cos_loss with the same signatures).
    Returns a flat float64 array."""
    kw = dict(case["kw"])
    fn = case["fn"]
    base, tc, params = kw.pop("base"), kw.pop("tc"), np.asarray(kw.pop("params"), dtype=np.float64)
    if fn == "carr_madan":
        out = impl.carr_madan(base, tc, params, kw["N"], kw["ada"], kw["alpha"], kw["S0"], kw["r"],
                              kw["T"], kw["is_put"])
    elif fn == "cos":
        out = impl.cos(base, tc, params, resolve_K(kw["K"]), kw["S0"], kw["r"], kw["T"], kw["numU"],
                       kw["kind"])
    elif fn == "fsts":
        out = impl.fsts(base, tc, params, kw["N"], kw["xmax"], kw["strike"], kw["r"], kw["T"],
                        kw["steps"], kw["american"], kw["payoff"], kw["kind"])
    elif fn == "objfn_cf":
        u = np.asarray(kw["u"])
        p2 = params.reshape(len(params), -1)
        # phiHat := exact log CF of set `phiHat_from_set` at 1 + i u (synthetic market, like
        # generateCharts.cpp:104-109 "exact" column), perturbed deterministically.
        ph = _phihat(base, tc, p2[kw["phiHat_from_set"]], u, kw["r"], kw["T"])
        out = impl.objfn_cf(base, tc, params, u, ph, kw["r"], kw["T"])
    elif fn == "cos_loss":
        K = resolve_K(kw["K"])
        p2 = params.reshape(len(params), -1)
        mkt = _c5_market(base, tc, p2[kw["mkt_from_set"]], K, kw["S0"], kw["r"], kw["T"], kw["numU"])
        w = c5_weights() if kw.get("w") is not None else None
        out = impl.cos_loss(base, tc, params, K, kw["S0"], kw["r"], kw["T"], kw["numU"], mkt, w)
    else:
        raise ValueError(fn)
    return np.asarray(out, dtype=np.float64).reshape(-1)


# "market" inputs for the two objective cases are produced by the PORT-independent closed forms
# below so that oracle and product see identical bytes.  They only need to be deterministic, not
# exact: they are inputs, not expectations.
_MKT_CACHE = {}


def _phihat(base, tc, p, u, r, T):
    key = ("ph", base, tc, tuple(p), tuple(u), r, T)
    if key not in _MKT_CACHE:
        # smooth synthetic target: a Gaussian log-CF with sigma = p[0] plus a small wiggle
        z = 1.0 + 1j * np.asarray(u)
        s = float(p[0])
        _MKT_CACHE[key] = ((-0.5 * s * s) * z + 0.5 * s * s * z * z) * T + 0.01 * np.sin(u) * (1 + 1j)
    return _MKT_CACHE[key]


def _c5_market(base, tc, p, K, S0, r, T, numU):
    key = ("mk", base, tc, tuple(p), tuple(K), S0, r, tuple(T), numU)
    if key not in _MKT_CACHE:
        s = float(p[0]) * 1.1
        mk = np.array([[bs_call(S0, math.exp(-r * t), k, s, t) for k in K] for t in T])
        _MKT_CACHE[key] = mk
    return _MKT_CACHE[key]
