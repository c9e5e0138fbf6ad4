// COS coefficients of one (parameter set, u_k) pair for every maturity (device code of
// cos_coeff_kernel; also compiled for the host by tests/hostemu, which runs the very same source
// against the oracle on a CPU): characteristic function -> option transform -> times V_k cp.
#pragma once
#include <type_traits>

#include "cf_models.cuh"

#ifdef __CUDA_ARCH__
#define FOP_LDG(p) __ldg(p)
#else
#define FOP_LDG(p) (*(p))
#endif

namespace fop {

struct CosCoeffArgs {
  int base, tc;
  int greek_mask;      // bit g set: emit the coefficients of transform g (0 price 1 delta 2 gamma 3 theta)
  size_t plane;        // doubles between the coefficient matrices of consecutive emitted transforms
  int P, M, K, Kp, np, spb;  // Kp = padded u count (row stride of C = 2 Kp); spb = sets per block
  double r;
  const double* params;  // [P][np]
  const double* T;       // [M]
  const int* npow;       // [M] steps d_m = n_m - n_{m-1} of T_m = n_m dt (< 0: restart), or nullptr (no common grid)
  double dt;
  const double* uk;      // [K]
  const double* vk;      // [K]
  double* C;             // [n_greeks][P*M][2 Kp]
  // quantised output of the int8 tensor-core contraction (cos_gemm_i8.cuh): six signed base-256
  // digit planes of the fixed-point coefficients, [P*M][6][2 Kp] bytes (the planes of one row are
  // adjacent: constant offsets from one pointer); a.vk then carries the fixed-point scale.  rowflag[row] is set when a coefficient of the row is not representable
  // (NaN / overflow): the contraction turns such rows into NaN like the fp64 path would.
  unsigned char* C8;
  unsigned char* rowflag;
  // (no further fields: with eight more bytes in this parameter block ptxas allocated the hot kernel 96
  // instead of 92 registers and it ran 1 % slower -- the two-set kernel's theta scale travels in vk[K])
};

// ---- fixed-point digit planes (int8 contraction) ---------------------------------------------
// value q (|q| < 2^47, integer) = sum_s d_s 256^s with SIGNED digits d_s in [-128, 127]: the bytes of
// (q + BIAS) ^ BIAS with BIAS = 0x808080808080 (adding 128 to every digit makes them the unsigned
// bytes of q + BIAS).
constexpr int Q8_SLICES = 6;
constexpr unsigned long long Q8_BIAS = 0x808080808080ull;
// x + Q8_MAGIC (one DADD, round to nearest even) leaves (rint(x) + BIAS) mod 2^48 in the low 48 mantissa
// bits: 2^52 + 2^51 + BIAS is exactly representable and its unit in the last place is 1.
constexpr double Q8_MAGIC = 6755399441055744.0 + 141289400074368.0;  // 1.5 * 2^52 + 0x808080808080
// Range test on the high word of x + Q8_MAGIC (1.5 * 2^52 has the high word 0x43380000, BIAS adds
// 0x8080, the integer its bits 32..47): accepted are (about) |rint(x)| < 2^46 -- twice the nominal
// range 2^45 of a coefficient; six signed digits reach 0.996 * 2^47.
constexpr unsigned Q8_HI_LO = 0x43380000u + 0x8080u - 0x4000u;
__host__ __device__ __forceinline__ bool q8_digits(double x, unsigned& lo, unsigned& hi) {
  const double t = x + Q8_MAGIC;
#ifdef __CUDA_ARCH__
  const unsigned long long b = (unsigned long long)__double_as_longlong(t);
#else
  unsigned long long b;
  __builtin_memcpy(&b, &t, 8);
#endif
  const unsigned h = (unsigned)(b >> 32);
  lo = (unsigned)b ^ 0x80808080u;
  hi = (h ^ 0x8080u) & 0xffffu;
  return (h - Q8_HI_LO) < 0x8000u;   // false: NaN or |x| >= 2^46
}
// row pointer of the quantised output: the (re, im) byte pair of digit plane 0; plane s lies Kp pairs
// further, the next maturity row 6 Kp pairs
struct Q8Row {
  unsigned short* p;
};
__host__ __device__ __forceinline__ void coeff_advance(double2*& c, int n, int Kp) { c += (size_t)n * Kp; }
__host__ __device__ __forceinline__ void coeff_advance(Q8Row& c, int n, int Kp) { c.p += (size_t)n * Q8_SLICES * Kp; }
__host__ __device__ __forceinline__ void coeff_store(const CosCoeffArgs&, double2* c, int i, int Kp, double x, double y) {
  c[(size_t)i * Kp] = make_double2(x, y);
}
__host__ __device__ __forceinline__ void coeff_store(const CosCoeffArgs& a, Q8Row c, int i, int Kp, double x, double y) {
  unsigned short* p = c.p + i * (Q8_SLICES * Kp);
  const int plane16 = Kp;
  unsigned xl, xh, yl, yh;
  const bool okx = q8_digits(x, xl, xh), oky = q8_digits(y, yl, yh);
  if (!(okx && oky)) {  // NaN or out of range: flag the row, store zeros
    const size_t off = (size_t)(p - reinterpret_cast<unsigned short*>(a.C8));
    a.rowflag[off / ((size_t)Q8_SLICES * Kp)] = 1;
    xl = xh = yl = yh = 0;
  }
#ifdef __CUDA_ARCH__
  p[0] = (unsigned short)__byte_perm(xl, yl, 0x0040);
  p[plane16] = (unsigned short)__byte_perm(xl, yl, 0x0051);
  p[2 * plane16] = (unsigned short)__byte_perm(xl, yl, 0x0062);
  p[3 * plane16] = (unsigned short)__byte_perm(xl, yl, 0x0073);
  p[4 * plane16] = (unsigned short)__byte_perm(xh, yh, 0x0040);
  p[5 * plane16] = (unsigned short)__byte_perm(xh, yh, 0x0051);
#else
  const unsigned long long ux = xl | ((unsigned long long)xh << 32), uy = yl | ((unsigned long long)yh << 32);
  for (int s = 0; s < Q8_SLICES; ++s)
    p[s * plane16] = (unsigned short)(((ux >> (8 * s)) & 0xff) | (((uy >> (8 * s)) & 0xff) << 8));
#endif
}

constexpr int COEFF_MAX_SPB = 64;

// Running power of E = exp(-delta dt) along the maturity list (maturities on the grid T_m = n_m dt,
// ascending; a.npow != nullptr): exp(-delta T_m) = E^{n_m} = R_{m-1} * E^{d_m} with the steps d_m =
// n_m - n_{m-1} >= 0 supplied by the host (cos_host_tables.h; a list that is not ascending does not
// use the grid).  E^{d} is kept while consecutive steps repeat, squared when the step doubles, and
// otherwise formed by binary powering (monthly lists: d = 1, 2, 3, 3, 3, 6, 6, 12 -> E, a square,
// one powering, two squares per (set, u) pair).  All branches are uniform over the grid.  This
// replaces one exp + sincos per maturity -- a quarter of the fp64 work of a Heston coefficient --
// by one complex multiplication.
struct PowState {
  cd R, D;     // R = E^{n_{m-1}}, D = E^{dprev}
  int dprev;
};
__host__ __device__ __forceinline__ cd csquare(cd z) { return mk(z.x * z.x - z.y * z.y, 2.0 * z.x * z.y); }
__host__ __device__ __forceinline__ cd pow_step(PowState& ps, cd E, int d) {
  if (d != ps.dprev) {
    if (d == 2 * ps.dprev) {
      ps.D = csquare(ps.D);
    } else if (d == 1) {
      ps.D = E;
    } else {
      cd acc = mk(1.0, 0.0), S = E;
      for (int b = d; b != 0; b >>= 1) {
        if (b & 1) acc = acc * S;
        S = csquare(S);
      }
      ps.D = acc;
    }
    ps.dprev = d;
  }
  ps.R = ps.R * ps.D;
  return ps.R;
}

// GRID: the maturities sit on a common grid (a.npow != nullptr; decided once per launch, outside the
// loops).  GMASK >= 0: the set of emitted transforms is a compile-time constant (1 = price only).
// FAST: CIR clock through the re-associated form (cir_logphi_fast) -- specialised kernels; the
// generic kernel keeps the term-by-term form of the reference (cf_log_cir_given_e).
template <int MODE, int N, bool GRID, int GMASK = -1, bool FAST = false, class CP = double2*>
__host__ __device__ __forceinline__ void coeff_group(const CosCoeffArgs& a, const ModelConsts& mc,
                                            const CfPre& pre, const CirFast& cf, cd E, PowState& ps,
                                            cd u, double v, CP crow, int Kp, int m) {
  // crow points at maturity m of this (set, u) pair (the caller advances it: no 64-bit row
  // multiplication per store)
  VD<N> Tm;
#pragma unroll
  for (int i = 0; i < N; ++i) Tm.v[i] = FOP_LDG(a.T + m + i);
  cx<VD<N>> lphi;
  if (MODE == CF_MODE_CIR && (GRID || FAST)) {
    cx<VD<N>> e;
    if (GRID) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const cd r = pow_step(ps, E, FOP_LDG(a.npow + m + i));
        e.x.v[i] = r.x;
        e.y.v[i] = r.y;
      }
    } else {
      e = cexp(mk(-pre.delta.x * Tm, -pre.delta.y * Tm));
    }
    if (FAST) {
      lphi = cir_logphi_fast(cf, Tm, e);
    } else {
      const cx<VD<N>> l = mk(pre.lin.x * Tm, pre.lin.y * Tm);
      lphi = cf_log_cir_given_e(mc, pre, Tm, l, e);
    }
  } else {
    lphi = cf_log_at_mode<MODE>(mc, pre, Tm);
  }
  const int gmask = GMASK >= 0 ? GMASK : a.greek_mask;
  if (gmask == 1) {  // price only: V_k cp folded into the exponential's scale
    const cx<VD<N>> t = cexp_scaled(lphi, v);
#pragma unroll
    for (int i = 0; i < N; ++i) coeff_store(a, crow, i, Kp, t.x.v[i], t.y.v[i]);
    return;
  }
  if constexpr (GMASK == 1) {
    // (price-only instantiation: nothing below is generated)
  } else if constexpr (std::is_same<CP, Q8Row>::value) {
    // quantised theta planes (GMASK = 8; price, delta and gamma are contracted from the price planes
    // above with their transform factor in the basis): v carries 2^45 / (4 + |r|), a bound of
    // |phi (log phi - r)| <= 1/e + pi + |r| for |phi| <= 1 (cos_gemm_i8.cuh)
    // GMASK = 9: both sets from one evaluation -- the theta planes lie a.plane bytes behind the price
    // planes (and their row flags chunk-rows behind, which the row index of coeff_store yields by itself)
    static_assert(GMASK == 1 || GMASK == 8 || GMASK == 9, "quantised output: price and / or theta planes");
    const cx<VD<N>> d = cexp(lphi);
    CP ct = crow;
    double vt = v;
    if (GMASK == 9) {
#pragma unroll
      for (int i = 0; i < N; ++i) coeff_store(a, crow, i, Kp, d.x.v[i] * v, d.y.v[i] * v);
      ct.p += a.plane / 2;
      vt = a.vk[a.K];   // the theta scale table [numU] follows the price scale table (capi_cos.cu)
    }
    const cx<VD<N>> t = option_transform(3, d, u, a.r);
#pragma unroll
    for (int i = 0; i < N; ++i) coeff_store(a, ct, i, Kp, t.x.v[i] * vt, t.y.v[i] * vt);
  } else {
    const cx<VD<N>> d = cexp(lphi);
    // one CF evaluation feeds every requested transform (price / delta / gamma / theta)
    double2* cg = crow;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (!(gmask >> g & 1)) continue;
      const cx<VD<N>> t = option_transform(g, d, u, a.r);
#pragma unroll
      for (int i = 0; i < N; ++i) cg[(size_t)i * Kp] = make_double2(t.x.v[i] * v, t.y.v[i] * v);
      cg += a.plane / 2;
    }
  }
}

// all maturities of one (set, u) pair, ILP at a time
template <int MODE, int ILP, bool GRID, int GMASK = -1, bool FAST = false, class CP = double2*>
__host__ __device__ __forceinline__ void coeff_maturities_g(const CosCoeffArgs& a, const ModelConsts& mc,
                                                 const CfPre& pre, cd u, double v, CP crow,
                                                 int Kp) {
  cd E = mk(1.0, 0.0);
  if (MODE == CF_MODE_CIR && GRID) E = cexp(mk(-pre.delta.x * a.dt, -pre.delta.y * a.dt));
  CirFast cf;
  if (MODE == CF_MODE_CIR && FAST) cf = cir_fast_prepare(mc, pre);
  PowState ps;
  ps.R = mk(1.0, 0.0);
  ps.D = mk(1.0, 0.0);
  ps.dprev = 0;
  const int M = a.M;
  int m = 0;
  for (; m + ILP <= M; m += ILP, coeff_advance(crow, ILP, Kp))
    coeff_group<MODE, ILP, GRID, GMASK, FAST, CP>(a, mc, pre, cf, E, ps, u, v, crow, Kp, m);
  for (; m < M; ++m, coeff_advance(crow, 1, Kp)) coeff_group<MODE, 1, GRID, GMASK, FAST, CP>(a, mc, pre, cf, E, ps, u, v, crow, Kp, m);
}
template <int MODE, int ILP, int GMASK = -1>
__host__ __device__ __forceinline__ void coeff_maturities(const CosCoeffArgs& a, const ModelConsts& mc,
                                                 const CfPre& pre, cd u, double v, double2* crow,
                                                 int Kp) {
  if (MODE == CF_MODE_CIR && a.npow) coeff_maturities_g<MODE, ILP, true, GMASK>(a, mc, pre, u, v, crow, Kp);
  else coeff_maturities_g<MODE, ILP, false, GMASK>(a, mc, pre, u, v, crow, Kp);
}

// Output map of the reference entry points (OptionPricing.h:378,414,439,464,490,515,541) as
//   out = (val - sub) * (disc * K * inv) + add      with per-kind uniform scalars
struct CosEpi { double sub, inv, add; };
__host__ __device__ inline CosEpi cos_epi_consts(int kind, double S0, double r) {
  switch (kind) {
    case 0: return {1.0, 1.0, S0};               // call price  (val-1) disc K + S0
    case 1: return {0.0, 1.0, 0.0};              // put price    val disc K
    case 2: return {0.0, 1.0 / S0, 1.0};         // call delta   val disc K / S0 + 1
    case 3: return {0.0, 1.0 / S0, 0.0};         // put delta
    case 4:
    case 5: return {0.0, 1.0 / (S0 * S0), 0.0};  // gamma
    case 6: return {r, 1.0, 0.0};                // call theta  (val - r) disc K
    default: return {0.0, 1.0, 0.0};             // put theta
  }
}

}  // namespace fop
