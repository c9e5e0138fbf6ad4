"""Mirror of namespace ``optionprice`` (reference OptionPricing.h) over the CUDA engine.

Same names and argument meaning as the reference; the ``cf`` argument is a ``models.CF`` descriptor
(evaluated on the device -- the fast path) or, as in the reference, ANY Python callable complex ->
complex: it is then sampled by the host on the grid the pricer defines and priced through the
``fop_*_samples`` entry points (everything after the evaluation runs on the GPU).  With a single
parameter set the functions return a 1-D array exactly like the reference's ``std::vector<double>``;
with P sets the leading axis is the set.  There is no CPU path: every function needs a CUDA device.
"""
from __future__ import annotations

import math

import numpy as np

from . import engine as E
from .models import CF

_default_engine = None


def default_engine() -> E.Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = E.Engine(0)
    return _default_engine


def _squeeze(out, cf: CF):
    return out[0] if cf.num_sets == 1 else out


# ---- grid helpers (OptionPricing.h:19-42, 84-138) -------------------------------------------
def getUDomain(uMax, du, index):
    myU = du * index
    return myU - 2.0 * uMax if myU > uMax else myU


def getMaxK(ada):
    return math.pi / ada


def getLambda(numSteps, b):
    return 2.0 * b / numSteps


def getStrikeUnderlying(xMin, xMax, K, numX):
    return default_engine().strike_underlying(xMin, xMax, K, numX)


def getFangOostStrike(xMin, xMax, S0, numX):
    return default_engine().fang_oost_strikes(xMin, xMax, S0, numX)


def getCarrMadanStrikes(ada, S0, numX):
    return default_engine().carr_madan_strikes(numX, ada, S0)


def getXFromK(S0, K):
    return np.log(S0 / np.asarray(K, dtype=np.float64))


# ---- Carr-Madan (OptionPricing.h:620-704) ---------------------------------------------------
def _sample(cf, z):
    """A Python callable on an array of complex arguments (vectorised if it accepts arrays)."""
    z = np.asarray(z, dtype=np.complex128)
    try:
        out = np.asarray(cf(z), dtype=np.complex128)
        if out.shape == z.shape:
            return out
    except Exception:
        pass
    return np.array([complex(cf(complex(v))) for v in z.ravel()]).reshape(z.shape)


def _carr_madan(numSteps, ada, args, is_put, engine):
    alpha, S0, r, T, cf = args if len(args) == 5 else (1.5,) + tuple(args)
    eng = engine or default_engine()
    if isinstance(cf, CF):
        return _squeeze(eng.carr_madan(cf.base, cf.time_change, cf.params, numSteps, ada, alpha, S0, r, T,
                                       is_put), cf)
    smp = _sample(cf, eng.carr_madan_u_grid(numSteps, ada, alpha))     # cf(alpha + 1 + i j ada)
    return eng.carr_madan_cf_samples(smp, numSteps, ada, alpha, S0, r, T, is_put)[0]


def CarrMadanCall(numSteps, ada, *args, engine=None):
    """CarrMadanCall(numSteps, ada, [alpha,] S0, r, T, cf); alpha defaults to 1.5 (:679)."""
    return _carr_madan(numSteps, ada, args, False, engine)


def CarrMadanPut(numSteps, ada, *args, engine=None):
    return _carr_madan(numSteps, ada, args, True, engine)


# ---- Fang-Oosterlee COS (OptionPricing.h:358-559) --------------------------------------------
def _cos(kind, S0, K, r, T, numUSteps, cf, engine):
    eng = engine or default_engine()
    if not isinstance(cf, CF):      # the reference's signature: cf is a callable for ONE maturity T
        assert np.ndim(T) == 0, "a callable characteristic function belongs to one maturity"
        smp = _sample(cf, 1j * eng.cos_u_grid(K, S0, numUSteps))
        return eng.cos_cf_samples(smp[None, None, :], K, S0, r, [T], numUSteps, kind)[0, 0]
    out = eng.cos(cf.base, cf.time_change, cf.params, K, S0, r, np.atleast_1d(T), numUSteps, kind)
    out = out[:, 0, :] if np.ndim(T) == 0 else out
    return _squeeze(out, cf)


def FangOostCallPrice(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.CALL_PRICE, S0, K, r, T, numUSteps, cf, engine)


def FangOostPutPrice(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.PUT_PRICE, S0, K, r, T, numUSteps, cf, engine)


def FangOostCallDelta(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.CALL_DELTA, S0, K, r, T, numUSteps, cf, engine)


def FangOostPutDelta(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.PUT_DELTA, S0, K, r, T, numUSteps, cf, engine)


def FangOostCallGamma(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.CALL_GAMMA, S0, K, r, T, numUSteps, cf, engine)


def FangOostPutGamma(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.PUT_GAMMA, S0, K, r, T, numUSteps, cf, engine)


def FangOostCallTheta(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.CALL_THETA, S0, K, r, T, numUSteps, cf, engine)


def FangOostPutTheta(S0, K, r, T, numUSteps, cf, engine=None):
    return _cos(E.PUT_THETA, S0, K, r, T, numUSteps, cf, engine)


# ---- FSTS (OptionPricing.h:185-281) ----------------------------------------------------------
def _fsts(kind, numSteps, xmax, r, T, strike, payoff, cf, engine, steps=1, american=False):
    """The reference passes getAssetPrice = K e^x and a payoff lambda (test.cpp:56-61); here the
    asset grid is fixed to K e^x and `payoff` is E.PAYOFF_CALL / E.PAYOFF_PUT."""
    eng = engine or default_engine()
    if not isinstance(cf, CF):
        # the reference's signature FSTS*(numSteps, xmax, r, T, getAssetPrice, payoff, cf): `strike` is
        # then the asset map and `payoff` the payoff callable (OptionPricing.h:185-281); cf is the CF
        # of one step of length T / steps
        get_asset, pay_fn = strike, payoff
        dx = 2.0 * xmax / (numSteps - 1.0)
        asset = np.array([get_asset(-xmax + dx * j) for j in range(numSteps)], dtype=np.float64)
        pay = np.array([pay_fn(a) for a in asset], dtype=np.float64)
        smp = _sample(cf, 1j * eng.fsts_u_grid(numSteps, xmax))
        return eng.fsts_samples(smp, pay, numSteps, xmax, r, T, steps=steps, american=american, kind=kind,
                                asset=asset)[0]
    return _squeeze(eng.fsts(cf.base, cf.time_change, cf.params, numSteps, xmax, strike, r, T,
                             steps, american, payoff, kind), cf)


def FSTSPrice(numSteps, xmax, r, T, strike, payoff, cf, engine=None, steps=1, american=False):
    return _fsts(E.FSTS_PRICE, numSteps, xmax, r, T, strike, payoff, cf, engine, steps, american)


def FSTSDelta(numSteps, xmax, r, T, strike, payoff, cf, engine=None):
    return _fsts(E.FSTS_DELTA, numSteps, xmax, r, T, strike, payoff, cf, engine)


def FSTSGamma(numSteps, xmax, r, T, strike, payoff, cf, engine=None):
    return _fsts(E.FSTS_GAMMA, numSteps, xmax, r, T, strike, payoff, cf, engine)


def FSTSTheta(numSteps, xmax, r, T, strike, payoff, cf, engine=None):
    return _fsts(E.FSTS_THETA, numSteps, xmax, r, T, strike, payoff, cf, engine)
