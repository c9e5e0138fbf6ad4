"""GPU tests of the int8 digit-plane COS contraction on tcgen05 (csrc/cos_gemm_i8.cuh) against the fp64
tensor-core contraction (FOP_GEMM_I8=0) and the oracle.

The int8 path is exact integer arithmetic on 46-bit fixed-point operands: its result differs from the
fp64 kernel's by the operand rounding only, ~2e-14 amax K_j per price (amax = max_k |V_k cp|, K_j the
strike) -- the bound asserted here is 1e-12 K_j, the parity bar against the reference stays
1e-8 + 1e-10 |ref| (OptionPricing.h:358-410)."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


def _both(engine, monkeypatch, fn):
    monkeypatch.delenv("FOP_GEMM_I8", raising=False)
    a = fn()
    monkeypatch.setenv("FOP_GEMM_I8", "0")
    b = fn()
    monkeypatch.delenv("FOP_GEMM_I8", raising=False)
    return a, b


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("P,T", [(1, [1.0]), (3, [0.3, 0.9, 1.7, 2.2, 3.1]), (129, list(CASES.C5_T)), (1000, [0.25, 0.5, 1.0])])
def test_i8_matches_fp64_contraction_and_oracle(kind, P, T, engine, best_oracle, monkeypatch):
    """Ragged row counts (1, 15, 1032, 3000 rows against 128-row tiles), call and put."""
    prm = CASES.heston_batch(P)
    K = CASES.c5_strikes()
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, kind))
    assert not np.array_equal(i8, f64), "the int8 contraction did not run (results identical to fp64)"
    assert np.all(np.abs(i8 - f64) <= 1e-12 * np.asarray(K)[None, None, :])
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, kind)
    assert_parity(i8, ref, what=f"int8 contraction kind {kind} P={P} M={len(T)}")


@pytest.mark.parametrize("model", ["bs", "merton", "cgmy", "vg", "merton_cir", "cgmy_cir_diffusion", "gauss_det_clock"])
def test_i8_every_model_family(model, engine, best_oracle, monkeypatch):
    cir = [1.2, 0.8, -0.4, 1.1]
    if model == "bs":
        base, tc, prm = CASES.GAUSS, CASES.TC_NONE, np.array([[0.2], [0.35]])
    elif model == "merton":
        base, tc, prm = CASES.MERTON, CASES.TC_NONE, CASES.merton_batch(9)
    elif model == "cgmy":
        base, tc, prm = CASES.CGMY, CASES.TC_NONE, CASES.cgmy_batch(9)
    elif model == "vg":
        base, tc, prm = 3, CASES.TC_NONE, np.array([[0.12, -0.14, 0.2], [0.25, -0.05, 0.5]])
    elif model == "merton_cir":
        base, tc, prm = CASES.MERTON, CASES.TC_CIR, np.hstack([CASES.merton_batch(9), np.tile(cir, (9, 1))])
    elif model == "cgmy_cir_diffusion":
        base, tc, prm = CASES.CGMY, CASES.TC_CIR_DIFFUSION, np.hstack([CASES.cgmy_batch(9), np.tile(cir, (9, 1))])
    else:
        base, tc, prm = CASES.GAUSS, CASES.TC_CIR, np.array([[0.3, 1.0, 0.0, -0.5, 1.0]])   # adaV = 0 (test.cpp:1187)
    K = np.concatenate([[5000.0], 100.0 * np.linspace(1.4, 0.6, 38), [3.0]])
    T = [0.25, 0.5, 1.0, 2.0]
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(base, tc, prm, K, 100.0, 0.03, T, 256, 0))
    assert not np.array_equal(i8, f64)
    assert np.all(np.abs(i8 - f64) <= 1e-12 * K[None, None, :])
    if model != "vg":
        ref = best_oracle.cos(base, tc, prm, K, 100.0, 0.03, T, 256, 0)
        assert_parity(i8, ref, what=f"int8 contraction, {model}")


@pytest.mark.parametrize("J", [2, 16, 71, 72, 73, 144, 145, 200, 1024])
def test_i8_strike_counts(J, engine, best_oracle, monkeypatch):
    """Strike counts around the 72-column block of the int8 kernel (seven 72-column level accumulators
    fill TMEM; more strikes = more blocks, blockIdx.y)."""
    prm = CASES.heston_batch(5)
    K = 100.0 * np.linspace(1.6, 0.4, J)
    T = [0.5, 1.0]
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 256, 0))
    assert engine.last_cos_contraction == "fp64-dmma"     # (the second of the two calls)
    assert not np.array_equal(i8, f64)
    assert np.abs(i8 - f64).max() <= 2e-10
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 256, 0)
    assert_parity(i8, ref, what=f"J={J}")


@pytest.mark.parametrize("M,expect", [(1, "int8-tcgen05"), (16, "int8-tcgen05"), (17, "fp64-dmma"), (40, "fp64-dmma")])
def test_i8_maturity_count_boundary(M, expect, engine, best_oracle):
    """Up to 16 maturities the market tables of a strike block fit next to the operand rings; longer
    lists fall back to the fp64 kernels (single- and multi-kind calls alike)."""
    prm = CASES.heston_batch(3)
    K = 100.0 * np.linspace(1.5, 0.5, 30)
    T = list(np.linspace(0.1, 2.0, M))
    got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, 1)
    assert engine.last_cos_contraction == expect
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, 1)
    assert_parity(got, ref, what=f"M={M}")
    monkey_kinds = engine.cos_greeks(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, [0, 6])
    assert engine.last_cos_contraction == expect and monkey_kinds.shape == (2, 3, M, 30)


def test_i8_loss_over_strike_blocks(engine, monkeypatch):
    """Row losses summed over several 72-column blocks (200 strikes = 3 blocks x 2 halves)."""
    prm = CASES.heston_batch(50)
    K = 100.0 * np.linspace(1.6, 0.4, 200)
    T = [0.25, 1.0, 2.0]
    mkt = np.array([[CASES.bs_call(100.0, 1.0, k, 0.3, t) for k in K] for t in T])
    w = np.abs(np.sin(np.arange(3 * 200))).reshape(3, 200)
    l8, l64 = _both(engine, monkeypatch, lambda: engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w))
    assert np.all(np.abs(l8 - l64) <= 1e-10 * np.abs(l64))


@pytest.mark.parametrize("numU", [1, 2, 33, 100, 1500])
def test_i8_u_counts_and_padding(numU, engine, best_oracle, monkeypatch):
    """numU off the 64-term chunk (zero-padded digit planes) and far from 256."""
    prm = CASES.heston_batch(4)
    K = 100.0 * np.linspace(1.5, 0.5, 21)
    T = [0.5, 1.0, 1.5]
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, numU, 0))
    assert np.abs(i8 - f64).max() <= 2e-10
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, numU, 0)
    assert_parity(i8, ref, what=f"numU={numU}")


def test_i8_loss_and_device_buffers(engine, monkeypatch):
    """fop_cos_loss through the int8 kernel: losses agree with the fp64 kernel to 1e-10 relative, with
    and without the price matrix, host and device outputs."""
    import torch
    P = 700
    prm = CASES.heston_batch(P)
    K, T = CASES.c5_strikes(), list(CASES.C5_T)
    mkt = np.array([[CASES.bs_call(100.0, 1.0, k, 0.25, t) for k in K] for t in T])
    w = CASES.c5_weights()
    (l8, p8), (l64, p64) = _both(engine, monkeypatch, lambda: engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w, want_prices=True))
    assert np.all(np.abs(l8 - l64) <= 1e-10 * np.abs(l64))
    assert np.all(np.abs(p8 - p64) <= 1e-12 * np.asarray(K)[None, None, :])
    dev = torch.device("cuda:0")
    loss = torch.zeros(P, dtype=torch.float64, device=dev)
    engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, torch.tensor(prm, device=dev), K, 100.0, 0.0, T, 256,
                    torch.tensor(mkt, device=dev), torch.tensor(w, device=dev), out=loss)
    engine.synchronize()
    assert np.array_equal(loss.cpu().numpy(), l8)      # deterministic: integer accumulation, fixed order


def test_i8_non_finite_parameters_give_nan_rows(engine, monkeypatch):
    """A parameter set whose characteristic function is NaN must come back as NaN (as from the fp64
    kernel and the reference), without disturbing its neighbours in the same 128-row tile."""
    prm = CASES.heston_batch(40)
    prm[7, 0] = np.nan
    prm[23, 2] = np.nan
    K = CASES.c5_strikes()
    T = [0.5, 1.0, 2.0]
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, 0))
    for p in (7, 23):
        assert np.all(np.isnan(i8[p])) and np.all(np.isnan(f64[p]))
    ok = [p for p in range(40) if p not in (7, 23)]
    assert np.all(np.isfinite(i8[ok]))
    assert np.all(np.abs(i8[ok] - f64[ok]) <= 1e-12 * np.asarray(K)[None, None, :])


def test_i8_wide_strike_range(engine, best_oracle, monkeypatch):
    """Strikes over six decades: the error of the fixed-point contraction scales with the strike
    (2e-14 amax K_j) and stays inside the parity bar."""
    prm = CASES.heston_batch(6)
    K = 100.0 * np.logspace(3, -3, 30)
    T = [0.5, 2.0]
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 512, 0))
    assert np.all(np.abs(i8 - f64) <= 1e-12 * K[None, None, :] + 1e-12)
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 512, 0)
    assert_parity(i8, ref, what="wide strike range")


def test_i8_chunked_batches_are_bit_identical(engine, monkeypatch):
    """FOP_COS_CHUNK_MIB=1 forces ~30 chunks of parameter sets (each with its own digit planes, row
    flags and tensor map): same bits as the single-chunk call, NaN rows included."""
    prm = CASES.heston_batch(900)
    prm[611, 1] = np.nan
    K, T = CASES.c5_strikes(), list(CASES.C5_T)
    mkt = np.array([[CASES.bs_call(100.0, 1.0, k, 0.25, t) for k in K] for t in T])
    w = CASES.c5_weights()
    monkeypatch.delenv("FOP_GEMM_I8", raising=False)
    one_l, one_p = engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w, want_prices=True)
    assert engine.last_cos_contraction == "int8-tcgen05"
    monkeypatch.setenv("FOP_COS_CHUNK_MIB", "1")
    many_l, many_p = engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w, want_prices=True)
    assert np.array_equal(one_p, many_p, equal_nan=True) and np.array_equal(one_l, many_l, equal_nan=True)
    assert np.all(np.isnan(one_p[611])) and np.isnan(one_l[611]) and np.isfinite(np.delete(one_l, 611)).all()


def test_i8_u_count_limit_falls_back(engine, best_oracle):
    """numU > 8192 (more than 16 384 terms per int32 accumulator) takes the fp64 kernels."""
    prm = CASES.heston_batch(2)
    K = 100.0 * np.linspace(1.4, 0.6, 9)
    got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, [1.0], 8192, 0)
    assert engine.last_cos_contraction == "int8-tcgen05"
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, [1.0], 8192, 0)
    assert_parity(got, ref, what="numU=8192")
    engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, [1.0], 8200, 0)
    assert engine.last_cos_contraction == "fp64-dmma"


@pytest.mark.parametrize("kind", [2, 3, 4, 5, 6, 7])
def test_i8_greeks(kind, engine, best_oracle, monkeypatch):
    """Call / put delta, gamma and theta (one kind per call) through the int8 kernel: the transform
    factor of delta and gamma (i u, -i u (1 - i u); OptionPricing.h:44-73) does not depend on the
    parameter set and lives in the basis planes -- they are contracted from the price planes --, theta's
    -(log phi - r) gets its own planes scaled by its bound 4 + |r|."""
    prm = CASES.heston_batch(33)
    K = CASES.c5_strikes()
    T = [0.25, 1.0, 2.0]
    i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.03, T, 256, kind))
    assert not np.array_equal(i8, f64), "the int8 contraction did not run"
    scale = np.asarray(K) / (100.0 if kind < 4 else 100.0**2 if kind < 6 else 0.2)   # output map: K / S0, K / S0^2, K
    assert np.all(np.abs(i8 - f64) <= 1e-10 * scale[None, None, :] + 1e-13)
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.03, T, 256, kind)
    assert_parity(i8, ref, what=f"int8 contraction kind {kind}")


def test_i8_tables_follow_the_requested_kind(engine, best_oracle):
    """price -> delta -> gamma -> theta -> price on one handle (and theta at another rate): the cached
    fixed-point tables are built per transform on demand."""
    prm = CASES.heston_batch(6)
    K = 100.0 * np.linspace(1.5, 0.5, 24)
    T = [0.5, 1.5]
    for kind, r in ((0, 0.02), (2, 0.02), (4, 0.02), (6, 0.02), (7, 0.07), (1, 0.02), (3, 0.02), (5, 0.02), (0, 0.02)):
        got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, r, T, 256, kind)
        assert engine.last_cos_contraction == "int8-tcgen05"
        ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, r, T, 256, kind)
        assert_parity(got, ref, what=f"kind {kind} after a table switch")
    # several kinds in one call: price, delta and gamma share ONE set of coefficient planes (phi itself;
    # their transform factors live in the basis planes) -- same bits as kind by kind; with theta in the
    # list one evaluation writes both sets of planes (another kernel instantiation: equal to the last
    # digit plane, not necessarily to the bit)
    kinds = [0, 2, 4, 1, 5]
    many = engine.cos_greeks(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, kinds)
    assert engine.last_cos_contraction == "int8-tcgen05"
    for i, kind in enumerate(kinds):
        one = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, kind)
        assert np.array_equal(many[i], one), f"kind {kind}"
    kinds = [0, 2, 4, 6, 7]
    many = engine.cos_greeks(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, kinds)
    for i, kind in enumerate(kinds):
        one = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, T, 256, kind)
        assert np.abs(many[i] - one).max() <= 1e-12 * max(1.0, np.abs(one).max()), f"kind {kind}"


@pytest.mark.parametrize("model", ["merton", "cgmy", "cgmy_cir_diffusion"])
def test_i8_greeks_other_models(model, engine, best_oracle, monkeypatch):
    cir = [1.2, 0.8, -0.4, 1.1]
    if model == "merton":
        base, tc, prm = CASES.MERTON, CASES.TC_NONE, CASES.merton_batch(5)
    elif model == "cgmy":
        base, tc, prm = CASES.CGMY, CASES.TC_NONE, CASES.cgmy_batch(5)
    else:
        base, tc, prm = CASES.CGMY, CASES.TC_CIR_DIFFUSION, np.hstack([CASES.cgmy_batch(5), np.tile(cir, (5, 1))])
    K = 100.0 * np.linspace(1.4, 0.6, 31)
    T = [0.5, 1.0]
    for kind in (2, 5, 6):
        i8, f64 = _both(engine, monkeypatch, lambda: engine.cos(base, tc, prm, K, 100.0, 0.03, T, 256, kind))
        assert not np.array_equal(i8, f64)
        ref = best_oracle.cos(base, tc, prm, K, 100.0, 0.03, T, 256, kind)
        assert_parity(i8, ref, what=f"int8 contraction, {model}, kind {kind}")
