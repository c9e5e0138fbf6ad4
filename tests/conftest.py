import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Parity bar of BASELINE.json north_star: |gpu - ref| <= 1e-8 + 1e-10 |ref| (fp64).
ATOL = 1e-8
RTOL = 1e-10


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def assert_parity(got, ref, atol=ATOL, rtol=RTOL, what=""):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    both_nan = np.isnan(got) & np.isnan(ref)
    err = np.abs(got - ref)
    tol = atol + rtol * np.abs(ref)
    bad = ~both_nan & ~(err <= tol)
    if bad.any():
        i = int(np.argmax(np.where(bad, err / tol, 0)))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{got.size} outside {atol:g}+{rtol:g}|ref|; worst at {i}: "
            f"got {got[i]!r} ref {ref[i]!r} err {err[i]:.3e} tol {tol[i]:.3e}")


def carr_madan_tolerance(ref, N, ada, alpha, atol=ATOL, rtol=RTOL, noise=5e-12):
    """Per-strike tolerance for Carr-Madan outputs.

    Output m is  S0 Re(X_m) e^{-alpha k_m} ada/(3 pi)  with k_m = -pi/ada + (2 pi/ada/N) m
    (OptionPricing.h:605): every rounding error of the spectrum/FFT is amplified by e^{-alpha k_m},
    up to e^{1.5*12.6} = 1.5e8 at the low-strike end.  The reference's own radix-2 FFT carries a
    twiddle-recurrence error of 1.4e-13 (2^10) .. 2.5e-12 (2^16) relative to max|X| (SURVEY.md
    section 0), so outside its test window [0.3N, 0.7N) (test.cpp:643-644) its output is that
    noise, not a price (values of magnitude 1e9 at m = 0), and can only be reproduced by a
    bit-identical FFT *and* libm.  Inside the window the north-star tolerance 1e-8 + 1e-10|ref| is
    applied unchanged; outside it the reference's own noise floor
    `noise * max_m|X_m-scale| * e^{-alpha k_m}` is added.
    """
    import numpy as _np
    ref = _np.asarray(ref, dtype=_np.float64).reshape(-1, N)
    b = _np.pi / ada
    k = -b + (2.0 * b / N) * _np.arange(N)
    damp = _np.exp(-alpha * k)
    lo, hi = int(N * .3), int(N * .7)
    tol = atol + rtol * _np.abs(ref)
    # un-damped magnitude of the spectrum, estimated from the trustworthy window
    scale = _np.max(_np.abs(ref[:, lo:hi]) / damp[lo:hi], axis=1, keepdims=True)
    extra = noise * scale * damp[None, :]
    extra[:, lo:hi] = 0.0
    return (tol + extra).reshape(-1)


def assert_parity_tol(got, ref, tol, what=""):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    err = np.abs(got - ref)
    bad = ~(np.isnan(got) & np.isnan(ref)) & ~(err <= tol)
    if bad.any():
        i = int(np.argmax(np.where(bad, err / tol, 0)))
        raise AssertionError(f"{what}: {int(bad.sum())}/{got.size} outside tolerance; worst at {i}: "
                             f"got {got[i]!r} ref {ref[i]!r} err {err[i]:.3e} tol {tol[i]:.3e}")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden_v1.json")) as f:
        g = json.load(f)
    return {c["name"]: c for c in g["cases"]}


def golden_expect(entry, out):
    """Select from `out` the entries the golden file kept; returns (got, expected)."""
    out = np.asarray(out, dtype=np.float64).reshape(-1)
    assert out.size == entry["n"], f"{entry['name']}: size {out.size} vs golden {entry['n']}"
    exp = np.asarray(entry["values"], dtype=np.float64)
    got = out if entry["idx"] is None else out[np.asarray(entry["idx"])]
    return got, exp


@pytest.fixture(scope="session")
def oracles():
    """Builds the port (always) and returns {'port': Oracle, 'reference': Oracle or None}."""
    from oracle import build as obuild
    from oracle import oracle as O
    obuild.build()
    out = {"port": O.load("port"), "reference": None}
    if O.available("reference"):
        out["reference"] = O.load("reference")
    return out


@pytest.fixture(scope="session")
def best_oracle(oracles):
    return oracles["reference"] or oracles["port"]
