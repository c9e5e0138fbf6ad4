"""GPU tests of the variance-gamma base model FOP_VG (BASELINE.json north_star: "CGMY/VG").  The
reference has no VG characteristic function, so parity is UNPINNED: the oracle's shim restates the
textbook form (Madan-Carr-Chang 1998), is itself checked against a 40-digit evaluation
(tests/test_hostemu.py), and every pricer is compared with the reference's pricers driven by that
CF."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import assert_parity, assert_parity_tol, carr_madan_tolerance

pytestmark = pytest.mark.gpu
VG = 3


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


def _cir(P):
    return CASES.heston_batch(P)[:, 1:]


@pytest.mark.parametrize("tc", [0, 1, 2])
def test_vg_cos_all_kinds(tc, engine, best_oracle):
    P = 33
    prm = CASES.vg_batch(P) if tc == 0 else np.hstack([CASES.vg_batch(P), _cir(P)])
    K, T = CASES.c5_strikes(), CASES.C5_T
    for kind in range(8):
        got = engine.cos(VG, tc, prm, K, 100.0, .05, T, 256, kind)
        ref = best_oracle.cos(VG, tc, prm, K, 100.0, .05, T, 256, kind)
        assert_parity(got, ref, what=f"VG COS tc={tc} kind={kind}")


def test_vg_carr_madan(engine, best_oracle):
    prm = CASES.vg_batch(9)
    for N in (1024, 16384):
        got = engine.carr_madan(VG, 0, prm, N, .25, 1.5, 100.0, .05, 1.0, True)
        ref = best_oracle.carr_madan(VG, 0, prm, N, .25, 1.5, 100.0, .05, 1.0, True)
        assert_parity_tol(got, ref, carr_madan_tolerance(ref, N, .25, 1.5), what=f"VG Carr-Madan N={N}")


def test_vg_fsts_single_step_and_american(engine, best_oracle):
    prm = CASES.vg_batch(5)
    for kind in range(4):
        got = engine.fsts(VG, 0, prm, 1024, 5.0, 100.0, .05, 1.0, payoff=1, kind=kind)
        ref = best_oracle.fsts(VG, 0, prm, 1024, 5.0, 100.0, .05, 1.0, payoff=1, kind=kind)
        assert_parity(got, ref, what=f"VG FSTS kind={kind}")
    got = engine.fsts(VG, 0, prm, 4096, 5.0, 100.0, .05, .5, steps=64, american=True, payoff=1)
    ref = best_oracle.fsts(VG, 0, prm, 4096, 5.0, 100.0, .05, .5, steps=64, american=True, payoff=1)
    assert_parity(got, ref, what="VG American put")


def test_vg_objective_and_calibration_recovers_parameters(engine, best_oracle):
    prm = CASES.vg_batch(64)
    u = np.linspace(-20, 20, 15)
    ph = np.array([best_oracle.cf(VG, 0, prm[0], 0.0, 1.0, complex(1.0, x))[1] for x in u])
    loss = engine.objfn_cf(VG, 0, prm, u, ph, 0.0, 1.0)
    np.testing.assert_allclose(loss, best_oracle.objfn_cf(VG, 0, prm, u, ph, 0.0, 1.0), rtol=1e-10, atol=1e-12)
    assert loss[0] < 1e-25
    p, fv, its, best = engine.calibrate_cf(VG, 0, [.05, -.4, .05], [.5, .2, .8], u, ph, 0.0, 1.0, nests=25,
                                           iterations=1500, tol=1e-12, seed=7, restarts=16)
    assert fv[best] < 1e-6
    np.testing.assert_allclose(p[best], prm[0], atol=2e-2)
