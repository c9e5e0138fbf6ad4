"""integrateFFT / integrateIFFT (fft.hpp:64-110).  CPU: the two oracle flavours
against each other, against golden vectors generated from the reference's own fft.hpp, and against
the analytic Fourier integral of a normal density.  GPU (-m gpu): fop_integrate_fft through the C
ABI against the oracle and the golden file, single-kernel and four-step sizes, host and device
buffers."""
import json
import os

import numpy as np
import pytest

from oracle.gen_golden_integrate import cases as golden_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_int():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "integrate_v1.json")))


def _inputs(name):
    return next(c for c in golden_inputs() if c["name"] == name)


def _check_golden(fn, golden_int, tol=1e-12):
    for g in golden_int["cases"]:
        c = _inputs(g["name"])
        y = fn(c["fn"], g["outputMin"], g["inputMin"], g["inputMax"], bool(g["inverse"]))
        ref = np.array(g["re"]) + 1j * np.array(g["im"])
        scale = 1.0 + np.max(np.abs(ref))
        assert np.max(np.abs(y[np.array(g["idx"])] - ref)) <= tol * scale, (g["name"], g["inverse"])


def test_port_matches_golden(oracles, golden_int):
    _check_golden(oracles["port"].integrate_fft, golden_int)


def test_reference_flavour_matches_golden_and_port(oracles, golden_int):
    ref = oracles["reference"]
    if ref is None:
        pytest.skip("reference flavour not built here")
    _check_golden(ref.integrate_fft, golden_int, tol=1e-15)
    rng = np.random.default_rng(3)
    f = rng.standard_normal((3, 512)) + 1j * rng.standard_normal((3, 512))
    for inv in (False, True):
        a = ref.integrate_fft(f, -2.0, -1.0, 5.0, inv)
        b = oracles["port"].integrate_fft(f, -2.0, -1.0, 5.0, inv)
        assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(a))


def test_density_integral_is_the_characteristic_function(oracles):
    """integrateFFT of a density: out_m = int f(x) e^{i (outputMin - m du) x} dx up to Simpson error
    (the forward transform's minus sign turns u_m into outputMin - m du inside the integral while
    the trailing phase uses u_m = outputMin + m du, exactly as coded in fft.hpp:70-81)."""
    N, lo, hi, mu, sg = 4096, -8.0, 8.0, 0.3, 0.8
    x = lo + (hi - lo) / (N - 1) * np.arange(N)
    f = np.exp(-0.5 * ((x - mu) / sg) ** 2) / (sg * np.sqrt(2 * np.pi))
    y = oracles["port"].integrate_fft(f + 0j, 0.0, lo, hi, False)
    du = ((N - 1) * 2 * np.pi / (hi - lo)) / N
    m = np.arange(40)
    v = -m * du                                    # frequency inside the integral
    inner = np.exp(1j * v * mu - 0.5 * (sg * v) ** 2) * np.exp(-1j * v * lo)   # sum starts at x_0
    expect = inner * np.exp(1j * (m * du) * lo)
    assert np.max(np.abs(y[:40] - expect)) <= 1e-9


# ------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


@pytest.mark.gpu
def test_gpu_matches_golden(engine, golden_int):
    _check_golden(engine.integrate_fft, golden_int)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [16, 64, 1024, 8192, 65536])
@pytest.mark.parametrize("inverse", [False, True])
def test_gpu_matches_oracle(engine, best_oracle, N, inverse):
    rng = np.random.default_rng(N + int(inverse))
    batch = 5 if N <= 8192 else 2
    f = rng.standard_normal((batch, N)) + 1j * rng.standard_normal((batch, N))
    got = engine.integrate_fft(f, -1.5, -2.0, 7.0, inverse)
    ref = best_oracle.integrate_fft(f, -1.5, -2.0, 7.0, inverse)
    # the reference's radix-2 FFT carries a twiddle-recurrence error
    assert np.max(np.abs(got - ref)) <= 5e-12 * np.max(np.abs(ref))


@pytest.mark.gpu
def test_gpu_device_buffers_and_aliasing(engine, best_oracle):
    import torch
    rng = np.random.default_rng(9)
    f = rng.standard_normal((7, 2048)) + 1j * rng.standard_normal((7, 2048))
    ref = best_oracle.integrate_fft(f, 0.0, 0.0, 3.0, False)
    t = torch.tensor(f, device="cuda:0")
    out = engine.integrate_fft(t, 0.0, 0.0, 3.0, False)
    engine.synchronize()
    assert np.max(np.abs(out.cpu().numpy() - ref)) <= 5e-12 * np.max(np.abs(ref))
    engine.integrate_fft(t, 0.0, 0.0, 3.0, False, out=t)      # in place
    engine.synchronize()
    assert np.max(np.abs(t.cpu().numpy() - ref)) <= 5e-12 * np.max(np.abs(ref))


@pytest.mark.gpu
def test_gpu_rejects_bad_arguments(engine):
    import fftoptionpricing_b200 as F
    with pytest.raises(F.FopError):
        engine.integrate_fft(np.zeros(100, dtype=np.complex128), 0.0, 0.0, 1.0)   # not a power of two
    with pytest.raises(F.FopError):
        engine.integrate_fft(np.zeros(64, dtype=np.complex128), 0.0, 1.0, 1.0)    # empty range


@pytest.mark.gpu
def test_python_mirror_of_fft_h(best_oracle):
    """fftoptionpricing_b200.fft: fft / ifft / integrateFFT with an ordinary callable integrand."""
    from fftoptionpricing_b200 import fft as FM
    rng = np.random.default_rng(5)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    assert np.max(np.abs(FM.ifft(FM.fft(x)) / 256 - x)) < 1e-13
    mu, sg, lo, hi, N = 0.3, 0.8, -8.0, 8.0, 1024
    dens = lambda t: np.exp(-0.5 * ((t - mu) / sg) ** 2) / (sg * np.sqrt(2 * np.pi))
    got = FM.integrateFFT(0.0, lo, hi, dens, N)
    grid = lo + (hi - lo) / (N - 1) * np.arange(N)
    ref = best_oracle.integrate_fft(dens(grid) + 0j, 0.0, lo, hi, False)
    assert np.max(np.abs(got - ref)) <= 5e-12 * np.max(np.abs(ref))
    doubled = FM.integrateFFT(0.0, lo, hi, dens, N, fnOutput=lambda u, v: 2.0 * v)
    assert np.max(np.abs(doubled - 2.0 * got)) == 0.0
