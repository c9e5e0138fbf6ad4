"""GPU parity tests of the two COS execution paths and of the maturity-grid shortcut.

* split path (default): cos_coeff_kernel + contraction -- every test of this file runs with BOTH
  contractions: the int8 digit-plane kernel on tcgen05 (default for call / put prices of up to 72
  strikes, cos_gemm_i8.cuh) and the fp64 tensor-core kernels (FOP_GEMM_I8=0);
* fused path (FOP_COS_PATH=fused): cos_fused_ws_kernel, warp-specialised producers/consumers;
* maturities on a common grid T_m = n_m dt use exp(-delta T_m) = exp(-delta dt)^{n_m}; off-grid
  maturities take one exp per maturity.
All against the oracle (reference FangOostCallPrice etc., OptionPricing.h:358-559) at the
north-star tolerance 1e-8 + 1e-10 |ref|."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import assert_parity

pytestmark = pytest.mark.gpu

ON_GRID = list(CASES.C5_T)                      # multiples of 1/12
OFF_GRID = [0.1, 0.2718281828, 0.7, 1.3]        # no common grid
WEEKLY = [k / 52.0 for k in (1, 2, 4, 13, 26, 52, 104)]
DESCENDING = [2.0, 1.0, 1.0, 0.25, 0.5, 3.0]     # on a grid, not sorted, one duplicate


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(autouse=True, params=["i8", "f64"])
def contraction(request, monkeypatch):
    if request.param == "f64":
        monkeypatch.setenv("FOP_GEMM_I8", "0")
    else:
        monkeypatch.delenv("FOP_GEMM_I8", raising=False)
    return request.param


def _strikes(J, S0=100.0):
    return np.concatenate([[S0 * 50.0], S0 * np.linspace(1.5, 0.5, J - 2), [S0 * 0.03]])


@pytest.mark.parametrize("path", ["split", "fused"])
@pytest.mark.parametrize("T", [ON_GRID, OFF_GRID, WEEKLY, [1.0], [0.5, 1.0, 1.5], DESCENDING],
                         ids=["monthly", "offgrid", "weekly", "single", "three", "unsorted"])
@pytest.mark.parametrize("J", [66, 33, 16])
def test_heston_paths_match_oracle(path, T, J, engine, best_oracle, monkeypatch):
    monkeypatch.setenv("FOP_COS_PATH", path)
    P = 37                                        # ragged: not a multiple of the 64-row tile
    prm = CASES.heston_batch(P)
    K = _strikes(J)
    for kind in (0, 1):                           # call / put price
        got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.03, T, 256, kind)
        ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.03, T, 256, kind)
        assert_parity(got, ref, what=f"{path} path, kind {kind}, M={len(T)}, J={J}")


@pytest.mark.parametrize("path", ["split", "fused"])
@pytest.mark.parametrize("model", ["merton", "cgmy", "merton_cir", "cgmy_cir_diffusion", "gauss_det_clock"])
def test_model_families_both_paths(path, model, engine, best_oracle, monkeypatch):
    monkeypatch.setenv("FOP_COS_PATH", path)
    K = _strikes(40)
    T = [0.25, 0.5, 1.0, 2.0]
    cir = [1.2, 0.8, -0.4, 1.1]
    if model == "merton":
        base, tc, prm = CASES.MERTON, CASES.TC_NONE, CASES.merton_batch(9)
    elif model == "cgmy":
        base, tc, prm = CASES.CGMY, CASES.TC_NONE, CASES.cgmy_batch(9)
    elif model == "merton_cir":
        base, tc = CASES.MERTON, CASES.TC_CIR
        prm = np.hstack([CASES.merton_batch(9), np.tile(cir, (9, 1))])
    elif model == "cgmy_cir_diffusion":
        base, tc = CASES.CGMY, CASES.TC_CIR_DIFFUSION
        prm = np.hstack([CASES.cgmy_batch(9), np.tile(cir, (9, 1))])
    else:  # adaV = 0: deterministic clock branch (test.cpp:1187)
        base, tc = CASES.GAUSS, CASES.TC_CIR
        prm = np.array([[0.3, 1.5, 0.0, -0.5, 1.0], [0.2, 0.7, 0.0, 0.3, 0.8]])
    got = engine.cos(base, tc, prm, K, 100.0, 0.02, T, 128, 0)
    ref = best_oracle.cos(base, tc, prm, K, 100.0, 0.02, T, 128, 0)
    assert_parity(got, ref, what=f"{path} path, {model}")


@pytest.mark.parametrize("path", ["split", "fused"])
def test_loss_and_greeks_both_paths(path, engine, best_oracle, monkeypatch):
    monkeypatch.setenv("FOP_COS_PATH", path)
    P = 21
    prm = CASES.heston_batch(P)
    K, T, w = CASES.c5_strikes(), CASES.C5_T, CASES.c5_weights()
    mkt = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm[:1], K, 100.0, 0.0, T, 256, 0)[0]
    got, prices = engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w,
                                  want_prices=True)
    ref = best_oracle.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, mkt, w)
    assert_parity(got, ref, what=f"{path} path fused loss")
    assert_parity(prices, best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, 0),
                  what=f"{path} path prices next to the loss")
    for kind in range(8):                         # every reference entry point (price/delta/gamma/theta)
        g = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 256, kind)
        r = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 256, kind)
        assert_parity(g, r, what=f"{path} path kind {kind}")


def test_paths_agree_at_bench_size(engine, monkeypatch):
    """Full config-5 shard slice: the two paths agree with each other to rounding, and the loss is
    reproducible run to run (fixed summation order)."""
    import torch
    P = 20000
    dev = torch.device("cuda:0")
    prm = torch.tensor(CASES.heston_batch(P), device=dev)
    K, T = CASES.c5_strikes(), CASES.C5_T
    outs = {}
    for path in ("split", "fused"):
        monkeypatch.setenv("FOP_COS_PATH", path)
        out = torch.empty((P, 8, 66), dtype=torch.float64, device=dev)
        engine.cos(0, 1, prm, K, 100.0, 0.0, T, 256, 0, out=out)
        engine.synchronize()
        outs[path] = out.cpu().numpy()
    assert np.all(np.isfinite(outs["split"]))
    err = np.abs(outs["split"] - outs["fused"])
    assert np.all(err <= 1e-9 + 1e-11 * np.abs(outs["split"])), err.max()


@pytest.mark.parametrize("path", ["split", "fused"])
def test_ragged_tiles_both_paths(path, engine, best_oracle, monkeypatch):
    """Tiles that are not full: M not dividing the 64-row tile (fused: SPT * M < 64), odd strike
    counts (scalar stores), numU not a multiple of the k-stage, one set, M > 32."""
    monkeypatch.setenv("FOP_COS_PATH", path)
    for P, M, J, numU in [(1, 1, 3, 64), (7, 5, 67, 100), (13, 3, 39, 256), (3, 11, 9, 17), (5, 33, 16, 40)]:
        prm = CASES.heston_batch(P)
        K = np.concatenate([[5000.0], np.linspace(180.0, 40.0, J - 2), [.3]])
        T = list(np.linspace(.1, 2.0, M))
        got = engine.cos(0, 1, prm, K, 100.0, .02, T, numU, 0)
        ref = best_oracle.cos(0, 1, prm, K, 100.0, .02, T, numU, 0)
        assert_parity(got, ref, what=f"{path} ragged P={P} M={M} J={J} numU={numU}")


@pytest.mark.parametrize("J", [33, 34, 65, 66])
@pytest.mark.parametrize("noex", ["", "1"])
def test_remainder_column_contraction(J, noex, engine, best_oracle, monkeypatch):
    """J = 8 NB + 1 or 2: the remainder columns run on DFMAs next to the tensor-core tile (default)
    or the strike block is padded (FOP_GEMM_NO_EX=1); prices, loss and every output kind."""
    monkeypatch.setenv("FOP_COS_PATH", "split")
    if noex:
        monkeypatch.setenv("FOP_GEMM_NO_EX", "1")
    else:
        monkeypatch.delenv("FOP_GEMM_NO_EX", raising=False)
    P = 19
    prm = CASES.heston_batch(P)
    K = _strikes(J)
    T = [0.25, 0.5, 1.0]
    for kind in (0, 1, 2, 4):
        got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, kind)
        ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, kind)
        assert_parity(got, ref, what=f"J={J} kind={kind} noex={noex!r}")
    mkt = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm[:1], K, 100.0, 0.01, T, 128, 0)[0] * 1.02
    w = np.ones((len(T), J))
    got = engine.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, mkt, w)
    ref = best_oracle.cos_loss(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.01, T, 128, mkt, w)
    assert_parity(got, ref, what=f"loss J={J} noex={noex!r}")


@pytest.mark.parametrize("numU", [1, 2, 1500, 4096])
def test_many_and_few_u_terms(numU, engine, best_oracle, monkeypatch):
    """numU far from the usual 256 (the reference takes any numUSteps, OptionPricing.h:362)."""
    monkeypatch.setenv("FOP_COS_PATH", "split")
    prm = CASES.heston_batch(3)
    K = _strikes(20)
    got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, [0.5, 1.0], numU, 0)
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.02, [0.5, 1.0], numU, 0)
    assert_parity(got, ref, what=f"numU={numU}")


@pytest.mark.parametrize("variant", ["generic", "shape0", "shape1", "shape2"])
@pytest.mark.parametrize("T", [ON_GRID, OFF_GRID, DESCENDING], ids=["monthly", "offgrid", "unsorted"])
def test_coefficient_kernel_variants_match_oracle(variant, T, engine, best_oracle, monkeypatch):
    """The generic coefficient kernel (term-by-term CIR clock, FOP_COEFF_GENERIC=1) and every CTA
    shape of the model-specialised one (re-associated clock, cos_coeff_spec.cuh) against the oracle;
    the specialised shapes must agree with each other bit for bit (same arithmetic, different
    launch geometry)."""
    monkeypatch.delenv("FOP_COS_PATH", raising=False)
    monkeypatch.delenv("FOP_COEFF_GENERIC", raising=False)
    monkeypatch.delenv("FOP_COEFF_SHAPE", raising=False)
    if variant == "generic":
        monkeypatch.setenv("FOP_COEFF_GENERIC", "1")
    else:
        monkeypatch.setenv("FOP_COEFF_SHAPE", variant[-1])
    prm = CASES.heston_batch(301)
    K = CASES.c5_strikes()
    got = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, 0)
    ref = best_oracle.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, 0)
    assert_parity(got, ref, what=f"coefficient kernel {variant}, M={len(T)}")
    if variant != "generic":
        monkeypatch.setenv("FOP_COEFF_SHAPE", "0")
        base = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, 0)
        assert np.array_equal(got, base)
    else:
        monkeypatch.delenv("FOP_COEFF_GENERIC", raising=False)
        spec = engine.cos(CASES.GAUSS, CASES.TC_CIR, prm, K, 100.0, 0.0, T, 256, 0)
        assert not np.array_equal(got, spec) and np.abs(got - spec).max() < 1e-8
