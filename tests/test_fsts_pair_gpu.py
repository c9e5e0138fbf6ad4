"""GPU parity tests of the paired FSTS kernel (two parameter sets per complex transform, used for
every batch of >= 2 sets) and of the unpaired one (FOP_FSTS_PAIR=0 / single set), against the
oracle (the reference's FSTSGeneric, OptionPricing.h:155-281, and the A.4 multi-step restatement)
for every grid size the single-CTA plan covers, odd batch sizes, all output kinds and the
early-exercise loop."""
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


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("N", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_single_step_all_sizes(N, pair, engine, best_oracle, monkeypatch):
    monkeypatch.setenv("FOP_FSTS_PAIR", pair)
    prm = CASES.merton_batch(5)                      # odd: the last slot carries one live set
    for payoff in (0, 1):
        got = engine.fsts(CASES.MERTON, 0, prm, N, 5.0, 50.0, 0.03, 1.0, payoff=payoff)
        ref = best_oracle.fsts(CASES.MERTON, 0, prm, N, 5.0, 50.0, 0.03, 1.0, payoff=payoff)
        assert_parity(got, ref, what=f"FSTS N={N} pair={pair} payoff={payoff}")


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_greeks_and_time_changed_models(kind, pair, engine, best_oracle, monkeypatch):
    monkeypatch.setenv("FOP_FSTS_PAIR", pair)
    for base, tc, prm in [(CASES.GAUSS, 0, np.array([[0.3], [0.2], [0.45]])),
                          (CASES.CGMY, 0, CASES.cgmy_batch(3)),
                          (CASES.GAUSS, 1, CASES.heston_batch(4))]:
        if kind == 3 and tc != 0:
            continue                                  # theta is only defined for Levy models
        got = engine.fsts(base, tc, prm, 1024, 5.0, 50.0, 0.03, 1.0, kind=kind)
        ref = best_oracle.fsts(base, tc, prm, 1024, 5.0, 50.0, 0.03, 1.0, kind=kind)
        assert_parity(got, ref, what=f"FSTS kind={kind} base={base} tc={tc} pair={pair}")


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("N,steps", [(256, 5), (1024, 16), (4096, 32)])
def test_american_put(N, steps, pair, engine, best_oracle, monkeypatch):
    monkeypatch.setenv("FOP_FSTS_PAIR", pair)
    prm = CASES.merton_batch(7)
    got = engine.fsts(CASES.MERTON, 0, prm, N, 7.5, 100.0, 0.05, 0.25, steps=steps, american=True, payoff=1)
    ref = best_oracle.fsts(CASES.MERTON, 0, prm, N, 7.5, 100.0, 0.05, 0.25, steps=steps, american=True, payoff=1)
    assert_parity(got, ref, what=f"American put N={N} steps={steps} pair={pair}")
    # early exercise: never below the payoff, never below the European value of the same grid
    S = CASES.strike_underlying(-7.5, 7.5, 100.0, N)
    assert np.all(got >= np.maximum(100.0 - S, 0.0)[None, :] - 1e-9)


def test_pairing_is_invisible(engine, monkeypatch):
    """The same set priced alone, as the first and as the second member of a pair gives the same
    numbers to rounding (the exchange couples the two members only through Hermitian symmetry)."""
    prm = CASES.merton_batch(4)
    monkeypatch.setenv("FOP_FSTS_PAIR", "1")
    both = engine.fsts(CASES.MERTON, 0, prm, 2048, 7.5, 100.0, 0.05, 0.25, steps=12, american=True, payoff=1)
    swapped = engine.fsts(CASES.MERTON, 0, prm[[1, 0, 3, 2]], 2048, 7.5, 100.0, 0.05, 0.25, steps=12,
                          american=True, payoff=1)
    alone = np.stack([engine.fsts(CASES.MERTON, 0, prm[i:i + 1], 2048, 7.5, 100.0, 0.05, 0.25, steps=12,
                                  american=True, payoff=1)[0] for i in range(4)])
    scale = 1.0 + np.abs(alone)
    assert np.max(np.abs(both - alone) / scale) < 1e-11
    assert np.max(np.abs(swapped[[1, 0, 3, 2]] - alone) / scale) < 1e-11


@pytest.mark.parametrize("N", [8192, 16384, 65536])
def test_large_grids_general_engine(N, engine, best_oracle):
    """N > 4096 runs through the general FFT engine (single-CTA transform up to 8192, four-step
    above): European single step with every output kind, and a short early-exercise loop."""
    prm = CASES.merton_batch(3)
    for kind in (0, 1, 2):
        got = engine.fsts(CASES.MERTON, 0, prm, N, 5.0, 50.0, 0.03, 1.0, kind=kind)
        ref = best_oracle.fsts(CASES.MERTON, 0, prm, N, 5.0, 50.0, 0.03, 1.0, kind=kind)
        # gamma multiplies the spectrum by -u(1-u) ~ (pi N / (2 xmax))^2 (2.6e7 at N = 16384) and
        # divides by asset^2 (down to 0.1): the reference's radix-2 FFT carries a twiddle-recurrence
        # error of ~1e-12 max|X| that this amplifies past 1e-8 at a few nodes
        atol = 1e-8 if (kind < 2 or N <= 8192) else 2e-7 * (N / 16384) ** 2
        assert_parity(got, ref, atol=atol, what=f"FSTS large N={N} kind={kind}")
    got = engine.fsts(CASES.MERTON, 0, prm, N, 7.5, 100.0, 0.05, 0.25, steps=4, american=True, payoff=1)
    ref = best_oracle.fsts(CASES.MERTON, 0, prm, N, 7.5, 100.0, 0.05, 0.25, steps=4, american=True, payoff=1)
    assert_parity(got, ref, what=f"American put large N={N}")
    hes = CASES.heston_batch(2)
    got = engine.fsts(CASES.GAUSS, 1, hes, N, 5.0, 50.0, 0.03, 1.0)
    ref = best_oracle.fsts(CASES.GAUSS, 1, hes, N, 5.0, 50.0, 0.03, 1.0)
    assert_parity(got, ref, what=f"FSTS large N={N} Heston")
