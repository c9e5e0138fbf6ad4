"""CPU tests of the DEVICE characteristic-function code: csrc/fmath.cuh, cplx.cuh, cf_models.cuh and
cos_coeff.cuh are __host__ __device__ source, tests/hostemu/emu.cu compiles the same source for the
host and these tests run it over parameter sweeps against the oracle -- branch cuts, special
cases and ill-conditioned corners of the hand-written elementary functions are checked here
without a GPU (the -m gpu sweeps in test_parity_sweep_gpu.py repeat them on the device at bench
scale).  The emulation is test infrastructure: the package never loads it."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import ATOL, RTOL


@pytest.fixture(scope="module")
def emu():
    from tests import hostemu
    try:
        return hostemu.HostEmu()
    except FileNotFoundError:
        pytest.skip("nvcc not available")


def _ratio(got, ref):
    P = got.shape[0]
    return (np.abs(got - ref) / (ATOL + RTOL * np.abs(ref))).reshape(P, -1).max(axis=1)


@pytest.mark.parametrize("tgrid", [True, False])
def test_emulated_heston_sweep_3000_sets_vs_oracle(emu, best_oracle, tgrid):
    prm = CASES.heston_batch(3000, seed=CASES.SEED + 11)
    K, T = CASES.c5_strikes(), CASES.C5_T
    ref = best_oracle.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0)
    got = emu.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0, use_tgrid=tgrid)
    r = _ratio(got, ref)
    assert r.max() <= 1.0, (int(np.argmax(r)), prm[int(np.argmax(r))], r.max())


@pytest.mark.parametrize("kind", [0, 1, 2, 4, 6])
def test_emulated_levy_models_all_kinds_vs_oracle(emu, best_oracle, kind):
    K = CASES.c5_strikes(500.0)
    for base, prm in [(2, CASES.cgmy_batch(64)), (1, CASES.merton_batch(64)), (0, np.linspace(.05, .6, 16)[:, None])]:
        ref = best_oracle.cos(base, 0, prm, K, 500.0, .04, [.25, 1.0, 2.5], 128, kind)
        got = emu.cos(base, 0, prm, K, 500.0, .04, [.25, 1.0, 2.5], 128, kind)
        r = _ratio(got, ref)
        assert r.max() <= 1.0, (base, kind, int(np.argmax(r)), r.max())


def test_emulated_time_changed_jump_models_vs_oracle(emu, best_oracle):
    K, T = CASES.c5_strikes(), CASES.C5_T
    h = CASES.heston_batch(48)[:, 1:]
    for base, tc, prm in [(1, 1, np.hstack([CASES.merton_batch(48), h])), (2, 1, np.hstack([CASES.cgmy_batch(48), h])),
                          (2, 2, np.hstack([CASES.cgmy_batch(48), h])), (0, 1, np.array([[.3, .4, 0.0, .3, 1.0]]))]:
        ref = best_oracle.cos(base, tc, prm, K, 100.0, .03, T, 128, 0)
        got = emu.cos(base, tc, prm, K, 100.0, .03, T, 128, 0)
        r = _ratio(got, ref)
        assert r.max() <= 1.0, (base, tc, int(np.argmax(r)), r.max())


def test_emulated_heston_box_edges(emu, best_oracle):
    """|rho| -> 1, v0Hat -> 0, slow / fast mean reversion at well-conditioned adaV: strict parity.
    adaV = 1e-4: the reference's own result is rounding noise (tests/mp_exact.py); the device code
    must be as close to the exact value as the reference is (factor 5)."""
    from tests import mp_exact as MP
    from tests.test_parity_sweep_gpu import _heston_edges
    prm = _heston_edges()
    K, T = CASES.c5_strikes(), CASES.C5_T
    ref = best_oracle.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0)
    got = emu.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0)
    r = _ratio(got, ref)
    well = prm[:, 2] > 1e-3
    assert r[well].max() <= 1.0
    worst = np.argsort(-r)[:3]
    for i in worst:
        ex = MP.cos_call_prices(0, 1, prm[i], K, 100.0, 0.0, T, 256)
        tol = ATOL + RTOL * np.abs(ex)
        r_ref = (np.abs(ref[i] - ex) / tol).max()
        r_emu = (np.abs(got[i] - ex) / tol).max()
        assert r_ref > 1.0 and r_emu <= 5.0 * r_ref, (prm[i], r_ref, r_emu)


def test_emulated_variance_gamma_vs_oracle_and_40_digit_reference(emu, best_oracle):
    """FOP_VG (north star "CGMY/VG"; absent from the reference, parity unpinned): the device code
    against the oracle's shim AND the shim against an independent 40-digit evaluation of the
    textbook CF."""
    from tests import mp_exact as MP
    K = CASES.c5_strikes()
    T = [.25, 1.0]
    prm = CASES.vg_batch(40)
    ref = best_oracle.cos(3, 0, prm, K, 100.0, .05, T, 256, 0)
    got = emu.cos(3, 0, prm, K, 100.0, .05, T, 256, 0)
    assert _ratio(got, ref).max() <= 1.0
    ex = MP.cos_call_prices(3, 0, prm[0], K, 100.0, .05, T, 256)
    assert (np.abs(ref[0] - ex) / (ATOL + RTOL * np.abs(ex))).max() <= 1.0
    h = CASES.heston_batch(12)[:, 1:]
    prm2 = np.hstack([CASES.vg_batch(12), h])
    for tc in (1, 2):
        ref = best_oracle.cos(3, tc, prm2, K, 100.0, .05, T, 128, 0)
        got = emu.cos(3, tc, prm2, K, 100.0, .05, T, 128, 0)
        assert _ratio(got, ref).max() <= 1.0, tc


def test_q8_digit_planes_round_to_nearest_and_flag_the_unrepresentable(emu):
    """The quantiser of the int8 contraction (cos_coeff.cuh::coeff_store(Q8Row), same source as the
    device): six signed base-256 digits reassemble to rint(x) (ties to even) for |x| < 2^46 -- twice the
    nominal coefficient range 2^45 -- and NaN / out-of-range values flag their row and store zeros."""
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-2.0**45, 2.0**45, 20000), rng.uniform(-300.0, 300.0, 2000),
                        [0.0, -0.0, 0.5, 1.5, 2.5, -0.5, -1.5, 127.5, 128.5, -128.5, 32767.5, 2.0**45, -2.0**45,
                         2.0**46 - 2.0**33, -(2.0**46 - 2.0**33)]])
    y = x[::-1].copy()
    qr, qi, flag, planes = emu.q8_store(x, y)
    assert not flag.any()
    assert np.array_equal(qr, np.rint(x).astype(np.int64))
    assert np.array_equal(qi, np.rint(y).astype(np.int64))
    bad = np.array([np.nan, np.inf, -np.inf, 2.0**46 + 2.0**33, -2.0**47, 1e300, 3.0])
    qr, qi, flag, planes = emu.q8_store(bad, np.zeros_like(bad))
    assert flag.tolist() == [1, 1, 1, 1, 1, 1, 0]
    assert not planes[:6].any() and qr[6] == 3
    qr, qi, flag, _ = emu.q8_store(np.array([1.0, 2.0]), np.array([np.nan, 5.0]))   # one bad component flags the pair
    assert flag.tolist() == [1, 0] and qr[1] == 2 and qi[1] == 5
