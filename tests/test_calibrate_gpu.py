"""GPU tests of fop_calibrate_cf, the on-device replacement of the reference's cuckoo::optimize call
(generateCharts.cpp:102-118).  There is no reference optimiser code to compare with (the library is
un-vendored and unpinned -- parity unpinned), so the tests are the properties the
reference's calibration driver relies on: the returned objective value is the objective at the
returned parameters (bit-identical to fop_objfn_cf, which IS pinned against getObjFn_arr), the
parameters respect the box, runs are reproducible, more iterations never hurt, and on a synthetic
market generated from known parameters (the recipe of generateCharts.cpp:43-118) the search drives
the objective to ~0 and recovers them."""
import numpy as np
import pytest

from tests import cases as CASES

pytestmark = pytest.mark.gpu

U = np.linspace(-20.0, 20.0, 15)                       # generateCharts.cpp uArray pattern
# generateCharts.cpp:217-223 (Heston boxes)
LO = np.array([0.0, 0.0, 0.0, -1.0, 0.0])
HI = np.array([0.6, 2.0, 4.0, 1.0, 2.0])


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


def _phihat(ora, base, tc, prm, r, T):
    return np.array([ora.cf(base, tc, prm, r, T, complex(1.0, u))[1] for u in U])


def test_value_is_objective_at_returned_parameters(engine, best_oracle):
    truth = CASES.heston0()
    ph = _phihat(best_oracle, 0, 1, truth, 0.0, 1.0)
    prm, fv, its, best = engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, nests=25, iterations=40,
                                             tol=0.0, seed=42, restarts=6)
    assert prm.shape == (6, 5) and np.all(its == 40)
    assert np.all(prm >= LO) and np.all(prm <= HI)
    again = engine.objfn_cf(0, 1, prm, U, ph, 0.0, 1.0)
    assert np.array_equal(again, fv)                   # same device code path -> same bits
    ref = best_oracle.objfn_cf(0, 1, prm, U, ph, 0.0, 1.0)
    assert np.all(np.abs(ref - fv) <= 1e-10 * (1 + np.abs(ref)))
    assert best == int(np.argmin(fv))


def test_reproducible_and_monotone(engine, best_oracle):
    ph = _phihat(best_oracle, 0, 1, CASES.heston0(), 0.0, 1.0)
    a = engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, iterations=30, tol=0.0, seed=7, restarts=3)
    b = engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, iterations=30, tol=0.0, seed=7, restarts=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, iterations=30, tol=0.0, seed=8, restarts=3)
    assert not np.array_equal(a[0], c[0])
    prev = None
    for iters in (0, 10, 100, 400):                    # same streams: a longer run extends a shorter one
        fv = engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, iterations=iters, tol=0.0, seed=7,
                                 restarts=3)[1]
        if prev is not None:
            assert np.all(fv <= prev)
        prev = fv


def test_recovers_heston_parameters(engine, best_oracle):
    """25 nests x 1500 iterations (generateCharts.cpp's nestSize / numSims), 148 restarts."""
    truth = np.array(CASES.heston0())
    ph = _phihat(best_oracle, 0, 1, truth, 0.0, 1.0)
    prm, fv, its, best = engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, nests=25, iterations=1500,
                                             tol=1e-14, seed=42, restarts=148)
    assert fv[best] < 1e-6, fv[best]
    # the objective pins the CF on u in [-20, 20]; compare CFs, and the well-identified parameters
    got = _phihat(best_oracle, 0, 1, prm[best], 0.0, 1.0)
    assert np.max(np.abs(got - ph)) < 1e-3
    assert abs(prm[best][0] - truth[0]) < 0.02         # sigma


def test_tolerance_stops_early_and_other_models(engine, best_oracle):
    truth = [0.2, 0.5, 0.05, 0.05]                     # Merton (generateCharts.cpp:324-329 boxes)
    lo, hi = [0.0, 0.0, -1.0, 0.0], [0.6, 2.0, 1.0, 2.0]
    ph = _phihat(best_oracle, 1, 0, truth, 0.04, 0.25)
    prm, fv, its, best = engine.calibrate_cf(1, 0, lo, hi, U, ph, 0.04, 0.25, iterations=1500,
                                             tol=1e-6, seed=1, restarts=32)
    assert fv[best] < 1e-6 and its[best] < 1500
    assert np.all(fv[its < 1500] < 1e-6)


def test_rejects_bad_arguments(engine):
    import fftoptionpricing_b200 as F
    ph = np.zeros(15, dtype=np.complex128)
    with pytest.raises(F.FopError):
        engine.calibrate_cf(0, 1, LO, HI, U, ph, 0.0, 1.0, nests=1)
    with pytest.raises(F.FopError):
        engine.calibrate_cf(0, 1, HI, LO, U, ph, 0.0, 1.0)
