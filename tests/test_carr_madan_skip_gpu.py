"""GPU test of the exact-zero shortcut of the Carr-Madan generator: points where the bound
|phi(a + i w)| <= phi(a) exp(-sigma^2 T w^2 / 2) proves that the reference's exp() underflows to
exactly 0 are not evaluated.  The output must be IDENTICAL (bit for bit, up to the sign of zero) to
the fully evaluated one (FOP_CM_NO_SKIP=1), for every model family, with and without a diffusion
part."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import assert_parity_tol, carr_madan_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("N", [1024, 8192, 65536])
@pytest.mark.parametrize("model", ["gauss", "merton", "cgmy", "cgmy_nosigma", "heston"])
def test_shortcut_is_bit_identical(N, model, engine, monkeypatch):
    if model == "gauss":
        base, tc, prm = CASES.GAUSS, 0, np.array([[0.3], [0.05], [0.6], [0.01]])
    elif model == "merton":
        base, tc, prm = CASES.MERTON, 0, CASES.merton_batch(4)
    elif model == "cgmy":
        base, tc, prm = CASES.CGMY, 0, CASES.cgmy_batch(4)
    elif model == "cgmy_nosigma":
        base, tc, prm = CASES.CGMY, 0, CASES.cgmy_batch(3).copy()
        prm[:, 0] = 0.0                                  # no diffusion: the bound never applies
    else:
        base, tc, prm = CASES.GAUSS, 1, CASES.heston_batch(3)   # time changed: never skipped
    args = (base, tc, prm, N, 0.25, 1.5, 500.0, 0.4, 0.25)
    monkeypatch.delenv("FOP_CM_NO_SKIP", raising=False)
    monkeypatch.setenv("FOP_CM_NO_PRUNE", "1")          # shortcut only, same transform
    fast = engine.carr_madan(*args)
    fast_put = engine.carr_madan(*args, is_put=True)
    monkeypatch.setenv("FOP_CM_NO_SKIP", "1")
    full = engine.carr_madan(*args)
    full_put = engine.carr_madan(*args, is_put=True)
    assert np.array_equal(fast, full), np.max(np.abs(fast - full))
    assert np.array_equal(fast_put, full_put)
    # default path: for N > 8192 the sets whose non-zero head fits 4096 / 8192 points go through
    # the input-pruned transform (a different summation order: equal up to the FFT rounding
    # noise that the undamping factor e^{-alpha k} amplifies, conftest.carr_madan_tolerance)
    monkeypatch.delenv("FOP_CM_NO_SKIP", raising=False)
    monkeypatch.delenv("FOP_CM_NO_PRUNE", raising=False)
    pruned = engine.carr_madan(*args)
    pruned_put = engine.carr_madan(*args, is_put=True)
    tol = carr_madan_tolerance(full, N, 0.25, 1.5)
    assert_parity_tol(pruned, full, tol, what=f"pruned vs full, {model}, N={N}")
    # the put adds S0 (K_m e^{-rT}/S0 - 1) (OptionPricing.h:652), up to 1e8 at the high-strike end: its
    # tolerance is relative to the put values, not to the calls'
    tol_put = np.maximum(tol, carr_madan_tolerance(full_put, N, 0.25, 1.5))
    assert_parity_tol(pruned_put, full_put, tol_put, what=f"pruned vs full (put), {model}, N={N}")
