"""CPU check of the ALGORITHM of the int8 digit-plane contraction (csrc/cos_gemm_i8.cuh): the
coefficient digits come from the device source compiled for the host (tests/hostemu: the very
`coeff_store(Q8Row)` the quantising coefficient kernels run), the rest -- basis digits, the 26 kept
plane pairs, the seven level sums, the three-group recombination -- is restated here in exact integer
arithmetic and compared with the plain fp64 contraction it replaces.  No GPU needed."""
import numpy as np
import pytest

from tests.hostemu import HostEmu

SLICES, LEVELS = 6, 7
A_SCALE, B_SCALE = 2.0**45, 2.0**46


@pytest.fixture(scope="module")
def emu():
    try:
        return HostEmu()
    except FileNotFoundError:
        pytest.skip("nvcc not available")


def digits(q):
    """Six signed base-256 digits of the integers q (least significant first)."""
    q = np.asarray(q, dtype=np.int64).copy()
    out = []
    for _ in range(SLICES):
        d = ((q + 128) % 256) - 128
        out.append(d)
        q = (q - d) // 256
    assert not q.any()
    return np.stack(out)            # [6][...]


def contract_q8(emu, phi, vk, basis, amax=None):
    """phi [rows][K] complex, vk [K] (row weights of the basis), basis [2K][J] -> (val [rows][J], level
    sums [7][rows][J])."""
    rows, K = phi.shape
    amax = np.abs(vk).max() if amax is None else amax
    # coefficient kernel: C' = (Re | Im) phi 2^45, quantised by the device code
    qr, qi, flag, _ = emu.q8_store((phi.real * A_SCALE).ravel(), (phi.imag * A_SCALE).ravel())
    assert not flag.any()
    qa = np.empty((rows, 2 * K), dtype=np.int64)
    qa[:, 0::2] = qr.reshape(rows, K)
    qa[:, 1::2] = qi.reshape(rows, K)
    # basis planes: Bas' = Bas V_k / amax 2^46 (cos_basis_q8_kernel)
    qb = np.rint(basis * (np.repeat(vk, 2) / amax)[:, None] * B_SCALE).astype(np.int64)
    a, b = digits(qa), digits(qb)                      # [6][rows][2K], [6][2K][J]
    S = np.zeros((LEVELS, rows, basis.shape[1]), dtype=np.int64)
    for s in range(SLICES):
        for t in range(SLICES):
            if s + t >= 4:
                S[s + t - 4] += a[s] @ b[t]            # exact int8 x int8 -> int32 tensor-core product
    assert np.abs(S).max() < 2**31                     # int32 accumulators
    lo = S[0] + S[1] * 256 + S[2] * 65536
    mid = S[3] + S[4] * 256 + S[5] * 65536
    top = S[6]
    assert max(np.abs(lo).max(), np.abs(mid).max()) < 2**48
    val = amax * (lo.astype(np.float64) * 2.0**-59 + mid.astype(np.float64) * 2.0**-35 + top.astype(np.float64) * 2.0**-11)
    return val, S


@pytest.mark.parametrize("K,J,rows", [(64, 5, 7), (256, 66, 16), (1000, 9, 3)])
def test_digit_plane_contraction_matches_fp64(emu, K, J, rows):
    rng = np.random.default_rng(K + J)
    u = np.pi / 7.0 * np.arange(K)
    x = np.linspace(0.0, 7.0, J)
    basis = np.empty((2 * K, J))
    basis[0::2] = np.cos(np.outer(u, x))
    basis[1::2] = -np.sin(np.outer(u, x))
    vk = (2.0 / 7.0) * np.cos(u * 0.3) / (1.0 + u * u) + 0.2 / (1.0 + np.arange(K)) * (-1.0) ** np.arange(K)
    vk[0] *= 0.5
    # |phi| <= 1, decaying like a characteristic function, random phases
    phi = np.exp(-0.02 * np.outer(rng.uniform(0.2, 3.0, rows), u * u)) * np.exp(1j * rng.uniform(-np.pi, np.pi, (rows, K)))
    coef = np.empty((rows, 2 * K))
    coef[:, 0::2] = phi.real * vk
    coef[:, 1::2] = phi.imag * vk
    ref = coef @ basis                                 # the fp64 contraction
    val, S = contract_q8(emu, phi, vk, basis)
    amax = np.abs(vk).max()
    # operand rounding: half a unit of 2^-45 resp. 2^-46 per term, weighted by the other operand (<= 1)
    bound = amax * (2.0**-46 * 2 * K + 2.0**-47 * 2 * K) + 2.0**-50
    assert np.abs(val - ref).max() <= bound
    # ... and far inside it in practice (errors are random): a few 1e-14 of amax
    assert np.abs(val - ref).max() <= 64 * 2.0**-46 * amax * np.sqrt(2 * K)


def test_level_sums_use_the_int32_range_safely(emu):
    """Worst-case digits (+-128 everywhere) at the host's limit of 16 384 terms stay inside int32."""
    worst = 6 * 128 * 128 * 16384
    assert worst < 2**31


@pytest.mark.parametrize("g", [1, 2])
def test_delta_and_gamma_live_in_the_basis_planes(emu, g):
    """The transform factor of delta (i u) and gamma (-(u^2 + i u)), OptionPricing.h:44-73, does not
    depend on phi: Re(phi c_k V_k e^{i u_k x_j}) = Re(phi) Re(z) - Im(phi) Im(z) with z = c_k e^{i u_k x_j},
    so the price planes (phi itself) contract against basis rows (Re z, -Im z) V_k / amax_g."""
    K, J, rows = 256, 40, 9
    rng = np.random.default_rng(g)
    u = np.pi / 6.0 * np.arange(K)
    x = np.linspace(0.0, 6.0, J)
    vk = (2.0 / 6.0) * np.cos(u * 0.4) / (1.0 + u * u) - 0.1 / (1.0 + np.arange(K))
    vk[0] *= 0.5
    phi = np.exp(-0.03 * np.outer(rng.uniform(0.2, 3.0, rows), u * u)) * np.exp(1j * rng.uniform(-np.pi, np.pi, (rows, K)))
    c = 1j * u if g == 1 else -(u * u + 1j * u)
    z = c[:, None] * np.exp(1j * np.outer(u, x))                       # [K][J]
    ref = np.einsum("rk,kj->rj", phi * vk, z).real                      # sum_k Re(phi c_k V_k e^{i u_k x_j})
    basis = np.empty((2 * K, J))
    basis[0::2] = z.real
    basis[1::2] = -z.imag
    f = np.abs(c)
    amax = (np.abs(vk) * f).max()
    fn = np.where(f > 0, f, 1.0)
    basis_unit = basis / np.repeat(fn, 2)[:, None]                      # rows of modulus <= 1 ...
    val, S = contract_q8(emu, phi, vk * f, basis_unit, amax=amax)       # ... times |c_k| V_k / amax_g
    assert np.abs(val - ref).max() <= 64 * 2.0**-46 * amax * np.sqrt(2 * K)
