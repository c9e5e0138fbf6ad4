// Branch-free fp64 elementary functions for the characteristic-function kernels.
//
// Why not libdevice exp/log/sincos/atan2: on sm_100a ptxas materialises every fp64 polynomial
// coefficient with two UMOVs (3 issue slots per DFMA -> the coefficient kernel was issue-bound at
// 55 % fp64-pipe utilisation, profiles/r1_cos_c5_split_v1_ncu.json), and the special-case branches
// of the library routines stop the scheduler from interleaving independent evaluations.  Here
//   * coefficients live in __constant__ memory (one LDCU.128 fetches two of them into uniform
//     registers that DFMA reads directly),
//   * every routine is straight-line code (selects instead of branches) and is a template over the
//     value type: double, or VD<N> = N independent values processed operation by operation.  The
//     warp scheduler issues in program order and a dependent DFMA can issue only every 8.4 cycles
//     (scripts/ub/ub_latency.cu: 1 chain x 4 warps/SMSP reaches 63 % of the fp64 pipe, 4 chains
//     90 %), and ptxas does not interleave independent call chains on its own
//     (profiles/r1_cos_fused_v1_sass_stalls.txt: Horner chains issued back to back) -- so the
//     interleaving of N evaluations is written into the source,
//   * accuracy is <= ~1.5 ulp on the domains the pricers use (tests/test_fmath_host.py measures
//     it against long double), far inside the 1e-10 parity budget.
// Polynomials: exp -- Chebyshev fit of (e^r - 1 - r)/r^2 on |r| <= ln2/2 (error 1.4e-18, generated with
// mpmath; scripts/gen_fmath_coeffs.py regenerates an equivalent degree-10 fit -- same error bound,
// coefficients equal to 2e-6 relative in the leading term and to rounding in the low ones -- and
// checks it against this table); sin/cos/log/atan -- the classical
// minimax sets of Sun's fdlibm (k_sin.c, k_cos.c, e_log.c, s_atan.c; public domain constants).
// Domains: exp any finite x (clamped to [-746, 710]); sincos |x| < 2^30; log x > 0 normal;
// atan2 finite.  NaNs propagate; sincos outside its domain returns NaN.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace fop {

#ifdef __CUDA_ARCH__
#define FM_TAB(name) name##_d
#else
#define FM_TAB(name) name##_h
#endif
// (Tried and dropped: a shared-memory copy of the tables read by volatile LDS at each use removes
// the MOV.SPILL / R2UR.FILL traffic ptxas generates around its uniform-register copies of the
// coefficients -- 17 % of the issued instructions -- but the kernels time the same within 1 %.)
#define FM_TABPTR const double*
// both copies are needed in a translation unit that compiles host and device passes
#define FM_TABLE(name, ...)                                   \
  static __constant__ double name##_d[] = {__VA_ARGS__};      \
  static const double name##_h[] = {__VA_ARGS__};

// q(r) = (e^r - 1 - r)/r^2, degree 10, highest power first
FM_TABLE(fm_exp_q, 2.0914717341716378e-09, 2.5105259538448494e-08, 2.755727357010214e-07,
         2.7557255297969643e-06, 2.4801587325605324e-05, 0.0001984126987490126,
         0.001388888888888373, 0.008333333333326112, 0.04166666666666667, 0.1666666666666667, 0.5)
// sin: r + r z (S1 + z (S2 + ... S6));  cos: 1 - z/2 + z^2 (C1 + z (C2 + ... C6)), z = r^2
FM_TABLE(fm_sin_s, 1.58969099521155010221e-10, -2.50507602534068634195e-08,
         2.75573137070700676789e-06, -1.98412698298579493134e-04, 8.33333333332248946124e-03,
         -1.66666666666666324348e-01)
FM_TABLE(fm_cos_c, -1.13596475577881948265e-11, 2.08757232129817482790e-09,
         -2.75573143513906633035e-07, 2.48015872894767294178e-05, -1.38888888888741095749e-03,
         4.16666666666666019037e-02)
// log: R(z) = z (Lg1 + z (Lg2 + ... Lg7)), z = s^2, s = f/(2+f)   (highest first)
FM_TABLE(fm_log_lg, 1.479819860511658591e-01, 1.531383769920937332e-01, 1.818357216161805012e-01,
         2.222219843214978396e-01, 2.857142874366239149e-01, 3.999999999940941908e-01,
         6.666666666666735130e-01)
// atan: t - t (z (aT0 + z (aT1 + ... aT10)))   (highest first), |t| <= tan(pi/8) < 7/16
FM_TABLE(fm_atan_t, 1.62858201153657823623e-02, -3.65315727442169155270e-02,
         4.97687799461593236017e-02, -5.83357013379057348645e-02, 6.66107313738753120669e-02,
         -7.69187620504482999495e-02, 9.09088713343650656196e-02, -1.11111104054623557880e-01,
         1.42857142725034663711e-01, -1.99999999998764832476e-01, 3.33333333333329318027e-01)

// scalar constants whose low 32 bits are non-zero (not encodable as DFMA immediates): kept in the
// constant bank as well so that no UMOV pair is needed to materialise them
FM_TABLE(fm_k, 1.4426950408889634,        // 0  log2(e)
         -6.93147180559945286e-01,        // 1  -ln2 hi
         -2.31904681384629956e-17,        // 2  -ln2 lo
         6.36619772367581382e-01,         // 3  2/pi
         -1.57079632679489656e+00,        // 4  -pi/2 hi
         -6.12323399573676604e-17,        // 5  -pi/2 mid
         1.49738490485916983e-33,         // 6  +|pi/2 lo|
         6.93147180369123816490e-01,      // 7  ln2_hi (fdlibm split)
         1.90821492927058770002e-10,      // 8  ln2_lo (fdlibm split)
         0.41421356237309503,             // 9  tan(pi/8)
         7.85398163397448279e-01,         // 10 pi/4 hi
         3.06161699786838302e-17,         // 11 pi/4 lo
         1.57079632679489656e+00,         // 12 pi/2 hi
         6.12323399573676604e-17,         // 13 pi/2 lo
         3.14159265358979312e+00,         // 14 pi hi
         1.22464679914735321e-16,         // 15 pi lo
         0.69314718055994530942)          // 16 ln2

#define FM_HD __host__ __device__ __forceinline__


// ---- N-wide value types: every operation is a loop over N independent lanes --------------------
template <int N> struct VD { double v[N]; };
template <int N> struct VI { int v[N]; };
template <int N> struct VB { bool v[N]; };

#define FM_LANES _Pragma("unroll") for (int i_ = 0; i_ < N; ++i_)
#define FM_VD_BINOP(op)                                                                             \
  template <int N> FM_HD VD<N> operator op(VD<N> a, VD<N> b) { VD<N> r; FM_LANES r.v[i_] = a.v[i_] op b.v[i_]; return r; } \
  template <int N> FM_HD VD<N> operator op(VD<N> a, double b) { VD<N> r; FM_LANES r.v[i_] = a.v[i_] op b; return r; }      \
  template <int N> FM_HD VD<N> operator op(double a, VD<N> b) { VD<N> r; FM_LANES r.v[i_] = a op b.v[i_]; return r; }
FM_VD_BINOP(+) FM_VD_BINOP(-) FM_VD_BINOP(*)
#undef FM_VD_BINOP
template <int N> FM_HD VD<N> operator-(VD<N> a) { VD<N> r; FM_LANES r.v[i_] = -a.v[i_]; return r; }
#define FM_VD_CMP(op)                                                                               \
  template <int N> FM_HD VB<N> operator op(VD<N> a, VD<N> b) { VB<N> r; FM_LANES r.v[i_] = a.v[i_] op b.v[i_]; return r; } \
  template <int N> FM_HD VB<N> operator op(VD<N> a, double b) { VB<N> r; FM_LANES r.v[i_] = a.v[i_] op b; return r; }
FM_VD_CMP(>) FM_VD_CMP(<) FM_VD_CMP(==) FM_VD_CMP(>=)
#undef FM_VD_CMP
#define FM_VI_BINOP(op)                                                                             \
  template <int N> FM_HD VI<N> operator op(VI<N> a, VI<N> b) { VI<N> r; FM_LANES r.v[i_] = a.v[i_] op b.v[i_]; return r; } \
  template <int N> FM_HD VI<N> operator op(VI<N> a, int b) { VI<N> r; FM_LANES r.v[i_] = a.v[i_] op b; return r; }         \
  template <int N> FM_HD VI<N> operator op(int a, VI<N> b) { VI<N> r; FM_LANES r.v[i_] = a op b.v[i_]; return r; }
FM_VI_BINOP(+) FM_VI_BINOP(-) FM_VI_BINOP(&) FM_VI_BINOP(|) FM_VI_BINOP(^) FM_VI_BINOP(>>) FM_VI_BINOP(<<)
#undef FM_VI_BINOP
#define FM_VI_CMP(op)                                                                               \
  template <int N> FM_HD VB<N> operator op(VI<N> a, int b) { VB<N> r; FM_LANES r.v[i_] = a.v[i_] op b; return r; }
FM_VI_CMP(>) FM_VI_CMP(<) FM_VI_CMP(!=)
#undef FM_VI_CMP
template <int N> FM_HD VB<N> operator&&(VB<N> a, VB<N> b) { VB<N> r; FM_LANES r.v[i_] = a.v[i_] && b.v[i_]; return r; }

// select (the scalar forms compile to the same FSEL / SEL as a ternary)
FM_HD double fm_sel(bool c, double a, double b) { return c ? a : b; }
FM_HD int fm_sel(bool c, int a, int b) { return c ? a : b; }
template <int N> FM_HD VD<N> fm_sel(VB<N> c, VD<N> a, VD<N> b) { VD<N> r; FM_LANES r.v[i_] = c.v[i_] ? a.v[i_] : b.v[i_]; return r; }
template <int N> FM_HD VD<N> fm_sel(VB<N> c, VD<N> a, double b) { VD<N> r; FM_LANES r.v[i_] = c.v[i_] ? a.v[i_] : b; return r; }
template <int N> FM_HD VD<N> fm_sel(VB<N> c, double a, VD<N> b) { VD<N> r; FM_LANES r.v[i_] = c.v[i_] ? a : b.v[i_]; return r; }
template <int N> FM_HD VD<N> fm_sel(VB<N> c, double a, double b) { VD<N> r; FM_LANES r.v[i_] = c.v[i_] ? a : b; return r; }
template <int N> FM_HD VI<N> fm_sel(VB<N> c, VI<N> a, VI<N> b) { VI<N> r; FM_LANES r.v[i_] = c.v[i_] ? a.v[i_] : b.v[i_]; return r; }
template <int N> FM_HD VI<N> fm_sel(VB<N> c, VI<N> a, int b) { VI<N> r; FM_LANES r.v[i_] = c.v[i_] ? a.v[i_] : b; return r; }
template <int N> FM_HD VI<N> fm_sel(VB<N> c, int a, VI<N> b) { VI<N> r; FM_LANES r.v[i_] = c.v[i_] ? a : b.v[i_]; return r; }
template <int N> FM_HD VI<N> fm_sel(VB<N> c, int a, int b) { VI<N> r; FM_LANES r.v[i_] = c.v[i_] ? a : b; return r; }

FM_HD int fm_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
FM_HD int fm_lo(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll);
#endif
}
FM_HD double fm_make(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  const long long b = ((long long)hi << 32) | (unsigned int)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}
FM_HD double fm_fma(double a, double b, double c) { return fma(a, b, c); }
FM_HD double fm_abs(double x) { return fabs(x); }
FM_HD double fm_i2d(int k) { return (double)k; }
FM_HD double fm_rcp_seed(double b) {  // >= 20 correct bits; b normal, nonzero
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  return r;
#else
  // host build (test harness): mimic MUFU.RCP64H -- only the high word of the operand is read and
  // only the high word of the result is produced (~20 correct bits), so that the Newton steps are
  // exercised with a seed as coarse as the device's
  long long bits; memcpy(&bits, &b, 8); bits &= 0xffffffff00000000ll; double bh; memcpy(&bh, &bits, 8);
  double r = 1.0 / bh; memcpy(&bits, &r, 8); bits &= 0xffffffff00000000ll; memcpy(&r, &bits, 8);
  return r;
#endif
}
FM_HD double fm_sqrt(double x) { return sqrt(x); }
FM_HD double fm_rsqrt_seed(double b) {  // >= 20 correct bits; b normal, positive
#ifdef __CUDA_ARCH__
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  return r;
#else
  // host build (test harness): as coarse as the hardware seed (see fm_rcp_seed)
  double r = 1.0 / sqrt(b);
  long long bits; memcpy(&bits, &r, 8); bits &= 0xfffffff800000000ll; memcpy(&r, &bits, 8);
  return r;
#endif
}

template <int N> FM_HD VI<N> fm_hi(VD<N> x) { VI<N> r; FM_LANES r.v[i_] = fm_hi(x.v[i_]); return r; }
template <int N> FM_HD VI<N> fm_lo(VD<N> x) { VI<N> r; FM_LANES r.v[i_] = fm_lo(x.v[i_]); return r; }
template <int N> FM_HD VD<N> fm_make(VI<N> hi, VI<N> lo) { VD<N> r; FM_LANES r.v[i_] = fm_make(hi.v[i_], lo.v[i_]); return r; }
template <int N> FM_HD VD<N> fm_make(VI<N> hi, int lo) { VD<N> r; FM_LANES r.v[i_] = fm_make(hi.v[i_], lo); return r; }
template <int N> FM_HD VD<N> fm_abs(VD<N> x) { VD<N> r; FM_LANES r.v[i_] = fabs(x.v[i_]); return r; }
template <int N> FM_HD VD<N> fm_i2d(VI<N> k) { VD<N> r; FM_LANES r.v[i_] = (double)k.v[i_]; return r; }
template <int N> FM_HD VD<N> fm_rcp_seed(VD<N> b) { VD<N> r; FM_LANES r.v[i_] = fm_rcp_seed(b.v[i_]); return r; }
template <int N> FM_HD VD<N> fm_sqrt(VD<N> x) { VD<N> r; FM_LANES r.v[i_] = sqrt(x.v[i_]); return r; }
#define FM_VD_FMA(TA, TB, TC, ea, eb, ec)                                                           \
  template <int N> FM_HD VD<N> fm_fma(TA a, TB b, TC c) { VD<N> r; FM_LANES r.v[i_] = fma(ea, eb, ec); return r; }
FM_VD_FMA(VD<N>, VD<N>, VD<N>, a.v[i_], b.v[i_], c.v[i_])
FM_VD_FMA(VD<N>, VD<N>, double, a.v[i_], b.v[i_], c)
FM_VD_FMA(VD<N>, double, VD<N>, a.v[i_], b, c.v[i_])
FM_VD_FMA(double, VD<N>, VD<N>, a, b.v[i_], c.v[i_])
FM_VD_FMA(VD<N>, double, double, a.v[i_], b, c)
FM_VD_FMA(double, VD<N>, double, a, b.v[i_], c)
#undef FM_VD_FMA
template <int N> FM_HD VD<N> fm_bcast(double x, VD<N>) { VD<N> r; FM_LANES r.v[i_] = x; return r; }
FM_HD double fm_bcast(double x, double) { return x; }

// reciprocal to full precision from the hardware seed (two Newton steps); b normal, nonzero
template <class T> FM_HD T fm_rcp(T b) {
  T r = fm_rcp_seed(b);
  T e = fm_fma(-b, r, 1.0);
  r = fm_fma(r, e, r);
  e = fm_fma(-b, r, 1.0);
  r = fm_fma(r, e, r);
  return r;
}
// 1/sqrt(x) and sqrt(x) for positive normal x, branch-free (no slow-path call as in the built-in sqrt:
// the CALL / BSSY scaffolding and its register moves cost more issue slots than the arithmetic).
// Seed 2^-20, two Newton steps (2^-80); sqrt = x * rsqrt with one residual correction (<= 1 ulp).
FM_HD double fm_rsqrt_pos(double x) {
  double y = fm_rsqrt_seed(x);
  const double h = 0.5 * x;
  y = fma(y, fma(-(h * y), y, 0.5), y);
  y = fma(y, fma(-(h * y), y, 0.5), y);
  return y;
}
FM_HD double fm_sqrt_pos(double x, double* rsqrt_out = nullptr) {
  const double y = fm_rsqrt_pos(x);
  const double s = x * y;
  if (rsqrt_out) *rsqrt_out = y;
  return fma(fma(-s, s, x), 0.5 * y, s);
}

// a / b to <= 1 ulp for normal-range operands.  One Newton step suffices for the reciprocal (seed
// 2^-20 -> 2^-40): the residual correction of the quotient squares that error again (2^-80).
template <class T> FM_HD T fm_div(T a, T b) {
  T r = fm_rcp_seed(b);
  r = fm_fma(r, fm_fma(-b, r, 1.0), r);
  const T q = a * r;
  return fm_fma(fm_fma(-b, q, a), r, q);
}

template <class T> FM_HD T fm_exp(T x) {
  FM_TABPTR K = FM_TAB(fm_k);
  // one magnitude clamp (keeps NaN: the comparison is false) so that k below fits an int
  x = fm_sel(fm_abs(x) > 1.0e6, fm_make((fm_hi(x) & 0x80000000) | 0x412e8480, 0), x);  // +-1e6
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const T t = fm_fma(x, K[0], MAGIC);
  auto k = fm_lo(t);
  const T kd = t - MAGIC;
  T r = fm_fma(kd, K[1], x);   // ln2 hi
  r = fm_fma(kd, K[2], r);   // ln2 lo
  FM_TABPTR c = FM_TAB(fm_exp_q);
  T q = fm_fma(r, c[0], c[1]);
#pragma unroll
  for (int i = 2; i <= 10; ++i) q = fm_fma(q, r, c[i]);
  const T p = fm_fma(r * r, q, r) + 1.0;
  // scale by 2^k in two normal steps; |k| clamped to 1100 -> exact 0 / inf outside the range
  k = fm_sel(k < -1100, -1100, fm_sel(k > 1100, 1100, k));
  const auto k1 = k >> 1, k2 = k - k1;
  return (p * fm_make((1023 + k1) << 20, 0)) * fm_make((1023 + k2) << 20, 0);
}

// sin and cos of x, |x| < 2^30
template <class T> FM_HD void fm_sincos(T x, T* sn, T* cs) {
  FM_TABPTR K = FM_TAB(fm_k);
  const double MAGIC = 6755399441055744.0;
  const T t = fm_fma(x, K[3], MAGIC);  // 2/pi
  const auto q = fm_lo(t);
  const T qd = t - MAGIC;
  T r = fm_fma(qd, K[4], x);   // pi/2 in three parts
  r = fm_fma(qd, K[5], r);
  r = fm_fma(qd, K[6], r);
  const T z = r * r;
  FM_TABPTR S = FM_TAB(fm_sin_s);
  FM_TABPTR C = FM_TAB(fm_cos_c);
  T ps = fm_fma(z, S[0], S[1]), pc = fm_fma(z, C[0], C[1]);
#pragma unroll
  for (int i = 2; i < 6; ++i) {
    ps = fm_fma(ps, z, S[i]);
    pc = fm_fma(pc, z, C[i]);
  }
  const T s0 = fm_fma(r * z, ps, r);
  const T c0 = fm_fma(z * z, pc, fm_fma(z, -0.5, 1.0));
  // quadrant: q&1 swaps, (q&2) negates sin, ((q+1)&2) negates cos -- sign flips are XORs of the
  // sign bit; outside the documented domain the result is NaN (loud), never silently wrong
  const auto ok = (fm_hi(x) & 0x7fffffff) < 0x41d00000;  // |x| < 2^30, false for NaN/inf
  const auto odd = (q & 1) != 0;
  const T ss = fm_sel(odd, c0, s0);
  const T cc = fm_sel(odd, s0, c0);
  const auto bad = fm_sel(ok, 0, 0x7ff80000);
  *sn = fm_make((fm_hi(ss) ^ ((q & 2) << 30)) | bad, fm_lo(ss));
  *cs = fm_make((fm_hi(cc) ^ (((q + 1) & 2) << 30)) | bad, fm_lo(cc));
}

// log(x), x > 0 and normal (callers scale; no special cases)
template <class T> FM_HD T fm_log_pos(T x) {
  FM_TABPTR K = FM_TAB(fm_k);
  auto hx = fm_hi(x);
  const auto lx = fm_lo(x);
  const auto k = (hx - 0x3fe6a09e) >> 20;  // x = 2^k m, m in [sqrt(1/2), sqrt(2))
  hx = hx - (k << 20);
  const T m = fm_make(hx, lx);
  const T f = m - 1.0;
  const T s = fm_div(f, 2.0 + f);
  const T z = s * s;
  FM_TABPTR L = FM_TAB(fm_log_lg);
  T R = fm_fma(z, L[0], L[1]);
#pragma unroll
  for (int i = 2; i < 7; ++i) R = fm_fma(R, z, L[i]);
  R = R * z;
  const T hfsq = 0.5 * f * f;
  const T dk = fm_i2d(k);
  // k ln2_hi - ((hfsq - (s (hfsq + R) + k ln2_lo)) - f)
  return fm_fma(dk, K[7], -((hfsq - fm_fma(s, hfsq + R, dk * K[8])) - f));
}

// atan2(y, x).  SCALE = true: any finite arguments ((0, 0) -> 0 or pi like the C library): the pair
// is first brought into [1, 2) by a power of two (tiny -- incl. subnormal, which rcp.approx
// flushes -- or huge operands; the ratio is unchanged).  SCALE = false: the caller guarantees
// 1e-290 < max(|x|, |y|) < 1e290 and saves the ten instructions of that step.
template <bool SCALE, class T> FM_HD T fm_atan2_impl(T y, T x) {
  FM_TABPTR K = FM_TAB(fm_k);
  T ax = fm_abs(x), ay = fm_abs(y);
  if (SCALE) {
    const T top = fm_sel(ax > ay, ax, ay);
    auto e = ((fm_hi(top) >> 20) & 0x7ff) - 1023;
    e = fm_sel(e < -1020, -1020, fm_sel(e > 1020, 1020, e));
    const T sc = fm_make((1023 - e) << 20, 0);  // 2^-e: the larger operand lands in [1, 2)
    ax = ax * sc;
    ay = ay * sc;
  }
  const auto ybig = ay > ax;
  const T mx = fm_sel(ybig, ay, ax), mn = fm_sel(ybig, ax, ay);
  const auto big = mn > K[9] * mx;  // tan(pi/8)
  const T num = fm_sel(big, mn - mx, mn);
  T den = fm_sel(big, mn + mx, mx);
  if (SCALE) den = fm_sel(den == 0.0, 1.0, den);  // atan2(0, 0): t = 0
  const T t = fm_div(num, den);
  const T z = t * t;
  FM_TABPTR A = FM_TAB(fm_atan_t);
  T p = fm_fma(z, A[0], A[1]);
#pragma unroll
  for (int i = 2; i < 11; ++i) p = fm_fma(p, z, A[i]);
  const T a0 = fm_fma(-t, z * p, t);  // atan(t), |t| <= tan(pi/8)
  // The octant fix-ups  a1 = pi/4 + a0 (big);  a2 = pi/2 - a1 (|y| > |x|);  a3 = pi - a2 (x < 0)
  // collapse to  m (pi/4) + s a0  with an integer m in 0..4 and a sign s: integer selects and one
  // fused multiply-add pair replace three select / subtract ladders on doubles.
  const auto neg = fm_hi(x) < 0;
  auto m = fm_sel(big, 1, 0);
  m = fm_sel(ybig, 2 - m, m);
  m = fm_sel(neg, 4 - m, m);
  // s = -1 iff exactly one of (|y| > |x|), (x < 0)
  const auto sflip = fm_sel(ybig, (int)0x80000000, 0) ^ (fm_hi(x) & (int)0x80000000);
  const T a0s = fm_make(fm_hi(a0) ^ sflip, fm_lo(a0));
  const T md = fm_i2d(m);
  const T a = fm_fma(md, K[10], fm_fma(md, K[11], a0s));  // m pi/4 (hi + lo) + s a0
  return fm_make(fm_hi(a) ^ (fm_hi(y) & 0x80000000), fm_lo(a));
}
template <class T> FM_HD T fm_atan2(T y, T x) { return fm_atan2_impl<true>(y, x); }
template <class T> FM_HD T fm_atan2_normal(T y, T x) { return fm_atan2_impl<false>(y, x); }

#undef FM_LANES

}  // namespace fop
