"""GPU tests of the caller-sampled entry points (fop_cos_cf_samples, fop_carr_madan_cf_samples,
fop_fsts_samples): the reference takes ANY callable as characteristic function / payoff
(OptionPricing.h:155-163, 325-333, 582-590); here the callable is evaluated by the host on the grid
the library defines and the samples are priced on the GPU.  The "callable" of these tests is the
reference's own CF (oracle.cf = the chfunctions shim called the way test.cpp calls it), so the
expected values are the reference pricers on the same model; a second set of cases uses a CF that is
NOT one of the built-in models (variance gamma written in Python)."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import assert_parity, assert_parity_tol, carr_madan_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


def _sample(ora, base, tc, p, r, T, zs):
    return np.array([ora.cf(base, tc, p, r, T, complex(z))[0] for z in zs])


@pytest.mark.parametrize("kind", [0, 1, 2, 4, 6])
def test_cos_from_sampled_cf_matches_reference(kind, engine, best_oracle):
    K = np.concatenate([[5000.0], np.linspace(180.0, 40.0, 31), [.3]])
    T, numU, S0, r = [.25, 1.0], 128, 100.0, .03
    u = engine.cos_u_grid(K, S0, numU)
    for base, tc, prm in [(0, 1, CASES.heston_batch(3)), (2, 0, CASES.cgmy_batch(3))]:
        smp = np.array([[_sample(best_oracle, base, tc, p, r, t, 1j * u) for t in T] for p in prm])
        got = engine.cos_cf_samples(smp, K, S0, r, T, numU, kind)
        ref = best_oracle.cos(base, tc, prm, K, S0, r, T, numU, kind)
        assert_parity(got, ref, what=f"cos_cf_samples kind={kind} model=({base},{tc})")
        dev = engine.cos(base, tc, prm, K, S0, r, T, numU, kind)
        assert np.abs(got - dev).max() < 1e-8


def test_carr_madan_from_sampled_cf_matches_reference(engine, best_oracle):
    N, ada, alpha, S0, r, T = 1024, .25, 1.5, 500.0, .4, .25
    z = engine.carr_madan_u_grid(N, ada, alpha)
    prm = CASES.cgmy_batch(3)
    smp = np.array([_sample(best_oracle, 2, 0, p, r, T, z) for p in prm])
    for put in (False, True):
        got = engine.carr_madan_cf_samples(smp, N, ada, alpha, S0, r, T, put)
        ref = best_oracle.carr_madan(2, 0, prm, N, ada, alpha, S0, r, T, put)
        assert_parity_tol(got, ref, carr_madan_tolerance(ref, N, ada, alpha), what=f"carr_madan_cf_samples put={put}")


@pytest.mark.parametrize("N", [1024, 16384])
def test_fsts_from_sampled_cf_and_payoff_matches_reference(N, engine, best_oracle):
    xmax, K, r, T = 5.0, 50.0, .05, 1.0
    w = engine.fsts_u_grid(N, xmax)
    S = CASES.strike_underlying(-xmax, xmax, K, N)
    prm = CASES.merton_batch(2)
    smp = np.array([_sample(best_oracle, 1, 0, p, r, T, 1j * w) for p in prm])
    for payoff_kind, pay in [(0, np.maximum(S - K, 0.0)), (1, np.maximum(K - S, 0.0))]:
        for kind in (0, 1, 2, 3):
            got = engine.fsts_samples(smp, pay, N, xmax, r, T, kind=kind, asset=S)
            ref = best_oracle.fsts(1, 0, prm, N, xmax, K, r, T, steps=1, american=False, payoff=payoff_kind, kind=kind)
            tol = 1e-8 + 1e-10 * np.abs(ref)
            if kind == 2 and N >= 16384:
                tol = tol + 2e-7      # gamma on long grids: reference FFT noise
            assert_parity_tol(got, ref, tol.reshape(-1), what=f"fsts_samples N={N} payoff={payoff_kind} kind={kind}")


def test_fsts_american_from_samples_matches_oracle_time_loop(engine, best_oracle):
    N, xmax, K, r, T, steps = 1024, 7.5, 100.0, .05, .25, 32
    w = engine.fsts_u_grid(N, xmax)
    S = CASES.strike_underlying(-xmax, xmax, K, N)
    prm = CASES.merton_batch(3)
    smp = np.array([_sample(best_oracle, 1, 0, p, r, T / steps, 1j * w) for p in prm])
    pay = np.maximum(K - S, 0.0)
    got = engine.fsts_samples(smp, pay, N, xmax, r, T, steps=steps, american=True)
    ref = best_oracle.fsts(1, 0, prm, N, xmax, K, r, T, steps=steps, american=True, payoff=1)
    assert_parity(got, ref, what="fsts_samples American put")
    # per-set payoffs: a different strike per set
    Ks = np.array([90.0, 100.0, 110.0])
    pays = np.maximum(Ks[:, None] - S[None, :], 0.0)
    got2 = engine.fsts_samples(smp, pays, N, xmax, r, T, steps=steps, american=True)
    assert np.array_equal(got2[1], got[1]) and not np.array_equal(got2[0], got[0])


def test_variance_gamma_callable_not_a_builtin_model(engine):
    """A CF outside the closed model set reaches the GPU: variance gamma, phi(u) = exp(u (r + w) T)
    (1 - u theta nu - sigma^2 nu u^2 / 2)^(-T/nu) at u = i w.  COS call prices must satisfy put-call
    parity with the sampled put prices and agree with a direct numpy COS sum."""
    sig, theta, nu, r, T, S0 = .12, -.14, .2, .1, 1.0, 100.0
    omega = np.log(1.0 - theta * nu - .5 * sig * sig * nu) / nu

    def vg(z):
        return np.exp(z * (r + omega) * T) * (1.0 - z * theta * nu - .5 * sig * sig * nu * z * z) ** (-T / nu)
    K = np.concatenate([[5000.0], np.linspace(130.0, 70.0, 13), [.3]])
    numU = 256
    u = engine.cos_u_grid(K, S0, numU)
    smp = vg(1j * u)[None, None, :]
    call = engine.cos_cf_samples(smp, K, S0, r, [T], numU, 0)[0, 0]
    put = engine.cos_cf_samples(smp, K, S0, r, [T], numU, 1)[0, 0]
    np.testing.assert_allclose(call - put, S0 - K * np.exp(-r * T), rtol=0, atol=1e-9)
    # direct evaluation of the COS sum in numpy
    x = np.log(S0 / K)
    a, b = x[0], x[-1]
    cp = 2.0 / (b - a)
    k = np.arange(numU)
    chi = (np.cos(u * (0 - a)) - np.exp(a) + u * np.sin(u * (0 - a))) / (1 + u * u)
    psi = np.where(k == 0, -a, np.sin(u * (0 - a)) / np.where(k == 0, 1.0, u))
    D = vg(1j * u) * cp * (psi - chi)
    D[0] *= .5
    val = (D[None, :] * np.exp(1j * u[None, :] * (x[:, None] - a))).real.sum(axis=1)
    np.testing.assert_allclose(put, val * np.exp(-r * T) * K, rtol=1e-10, atol=1e-9)
    assert 0 < call[7] < S0      # at-the-money call is a sane number (about 10.98 for these parameters)


def test_python_mirror_takes_the_references_callables(engine):
    """fftoptionpricing_b200.pricing with the reference's own call shapes: lambdas for the
    characteristic function, the payoff and the asset map (test.cpp:47-79, 342-367, 619-648) against
    the on-device Black-Scholes model."""
    from fftoptionpricing_b200 import models, pricing
    r, sig, T, S0, K, xmax, numX = .05, .3, 1.0, 50.0, 50.0, 5.0, 1024

    def bscf(u):
        return np.exp(r * T * u + sig * sig * .5 * T * (u * u - u))
    bs = models.black_scholes(sig)
    a = pricing.FSTSPrice(numX, xmax, r, T, lambda x: K * np.exp(x), lambda s: max(s - K, 0.0), bscf, engine=engine)
    b = pricing.FSTSPrice(numX, xmax, r, T, K, 0, bs, engine=engine)
    lo, hi = int(numX * .3), int(numX * .7)
    assert np.abs(a - b)[lo:hi].max() < 1e-9
    a = pricing.CarrMadanCall(numX, .25, S0, r, T, bscf, engine=engine)
    b = pricing.CarrMadanCall(numX, .25, S0, r, T, bs, engine=engine)
    assert np.abs(a - b)[lo:hi].max() < 1e-9
    Ks = [7500, 60, 50, 40, .3]
    for fn in (pricing.FangOostCallPrice, pricing.FangOostPutPrice, pricing.FangOostCallDelta, pricing.FangOostCallGamma):
        a = fn(S0, Ks, r, T, 64, bscf, engine=engine)
        b = fn(S0, Ks, r, T, 64, bs, engine=engine)
        assert np.abs(a - b)[1:4].max() < 1e-10
