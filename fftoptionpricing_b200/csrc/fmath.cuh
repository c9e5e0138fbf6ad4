// Branch-free fp64 elementary functions for the characteristic-function kernels.
//
// Why not libdevice exp/log/sincos/atan2: on sm_100a ptxas materialises every fp64 polynomial
// coefficient with two UMOVs (3 issue slots per DFMA -> the coefficient kernel was issue-bound at
// 55 % fp64-pipe utilisation, profiles/r1_cos_c5_split_v1_ncu.json), and the special-case branches
// of the library routines stop the scheduler from interleaving independent evaluations.  Here
//   * coefficients live in __constant__ memory (one LDCU.128 fetches two of them into uniform
//     registers that DFMA reads directly),
//   * every routine is straight-line code (selects instead of branches), so ILP copies of a
//     characteristic-function evaluation interleave in one basic block,
//   * accuracy is <= ~1.5 ulp on the domains the pricers use (tests/test_fmath_host.py measures
//     it against long double), far inside the 1e-10 parity budget.
// Polynomials: exp -- Chebyshev-economised fit of (e^r - 1 - r)/r^2 on |r| <= ln2/2 (error 1.4e-18,
// generated with mpmath, see scripts/gen_fmath_coeffs.py); sin/cos/log/atan -- the classical
// minimax sets of Sun's fdlibm (k_sin.c, k_cos.c, e_log.c, s_atan.c; public domain constants).
// Domains: exp any finite x (clamped to [-746, 710]); sincos |x| < 2^30; log x > 0 normal;
// atan2 finite.  NaNs propagate; sincos outside its domain returns NaN.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace fop {

#ifdef __CUDA_ARCH__
#define FM_TAB(name) name##_d
#define FM_DECL(name, ...) static __constant__ double name##_d[] = {__VA_ARGS__};
#else
#define FM_TAB(name) name##_h
#define FM_DECL(name, ...) static const double name##_h[] = {__VA_ARGS__};
#endif
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

// approximate reciprocal (>= 20 bits) refined to full precision; b normal, nonzero
FM_HD double fm_rcp(double b) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
  double r = 1.0 / b;  // host build is a test harness only
#endif
  double e = fm_fma(-b, r, 1.0);
  r = fm_fma(r, e, r);
  e = fm_fma(-b, r, 1.0);
  r = fm_fma(r, e, r);
  return r;
}
// a / b to <= 1 ulp for normal-range operands
FM_HD double fm_div(double a, double b) {
  const double r = fm_rcp(b);
  const double q = a * r;
  return fm_fma(fm_fma(-b, q, a), r, q);
}

FM_HD double fm_exp(double x) {
  const double* K = FM_TAB(fm_k);
  // one magnitude clamp (keeps NaN: the comparison is false) so that k below fits an int
  x = fabs(x) > 1.0e6 ? copysign(1.0e6, x) : x;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fm_fma(x, K[0], MAGIC);
  int k = fm_lo(t);
  const double kd = t - MAGIC;
  double r = fm_fma(kd, K[1], x);   // ln2 hi
  r = fm_fma(kd, K[2], r);   // ln2 lo
  const double* c = FM_TAB(fm_exp_q);
  double q = c[0];
#pragma unroll
  for (int i = 1; i <= 10; ++i) q = fm_fma(q, r, c[i]);
  const double p = fm_fma(r * r, q, r) + 1.0;
  // scale by 2^k in two normal steps; |k| clamped to 1100 -> exact 0 / inf outside the range
  k = k < -1100 ? -1100 : (k > 1100 ? 1100 : k);
  const int k1 = k >> 1, k2 = k - k1;
  return (p * fm_make((1023 + k1) << 20, 0)) * fm_make((1023 + k2) << 20, 0);
}

// sin and cos of x, |x| < 2^30
FM_HD void fm_sincos(double x, double* sn, double* cs) {
  const double* K = FM_TAB(fm_k);
  const double MAGIC = 6755399441055744.0;
  const double t = fm_fma(x, K[3], MAGIC);  // 2/pi
  const int q = fm_lo(t);
  const double qd = t - MAGIC;
  double r = fm_fma(qd, K[4], x);   // pi/2 in three parts
  r = fm_fma(qd, K[5], r);
  r = fm_fma(qd, K[6], r);
  const double z = r * r;
  const double* S = FM_TAB(fm_sin_s);
  const double* C = FM_TAB(fm_cos_c);
  double ps = S[0], pc = C[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) {
    ps = fm_fma(ps, z, S[i]);
    pc = fm_fma(pc, z, C[i]);
  }
  const double s0 = fm_fma(r * z, ps, r);
  const double c0 = fm_fma(z * z, pc, fm_fma(-0.5, z, 1.0));
  // quadrant: q&1 swaps, (q&2) negates sin, ((q+1)&2) negates cos -- sign flips are XORs of the
  // sign bit; outside the documented domain the result is NaN (loud), never silently wrong
  const bool ok = (fm_hi(x) & 0x7fffffff) < 0x41d00000;  // |x| < 2^30, false for NaN/inf
  const double ss = (q & 1) ? c0 : s0;
  const double cc = (q & 1) ? s0 : c0;
  const int bad = ok ? 0 : 0x7ff80000;
  *sn = fm_make((fm_hi(ss) ^ ((q & 2) << 30)) | bad, fm_lo(ss));
  *cs = fm_make((fm_hi(cc) ^ (((q + 1) & 2) << 30)) | bad, fm_lo(cc));
}

// log(x), x > 0 and normal (callers scale; no special cases)
FM_HD double fm_log_pos(double x) {
  const double* K = FM_TAB(fm_k);
  int hx = fm_hi(x);
  const int lx = fm_lo(x);
  const int k = (hx - 0x3fe6a09e) >> 20;  // x = 2^k m, m in [sqrt(1/2), sqrt(2))
  hx -= k << 20;
  const double m = fm_make(hx, lx);
  const double f = m - 1.0;
  const double s = fm_div(f, 2.0 + f);
  const double z = s * s;
  const double* L = FM_TAB(fm_log_lg);
  double R = L[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) R = fm_fma(R, z, L[i]);
  R *= z;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  // k ln2_hi - ((hfsq - (s (hfsq + R) + k ln2_lo)) - f)
  return fm_fma(dk, K[7], -((hfsq - fm_fma(s, hfsq + R, dk * K[8])) - f));
}

// atan2(y, x), finite arguments; (0, 0) -> 0 or pi like the C library
FM_HD double fm_atan2(double y, double x) {
  const double* K = FM_TAB(fm_k);
  double ax = fabs(x), ay = fabs(y);
  // bring tiny (incl. subnormal: rcp.approx flushes them) or huge pairs into the safe range;
  // powers of two, so the ratio is unchanged
  const double top = ax > ay ? ax : ay;
  int e = ((fm_hi(top) >> 20) & 0x7ff) - 1023;
  e = e < -1020 ? -1020 : (e > 1020 ? 1020 : e);
  const double sc = fm_make((1023 - e) << 20, 0);  // 2^-e: the larger operand lands in [1, 2)
  ax *= sc;
  ay *= sc;
  const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
  const bool big = mn > K[9] * mx;  // tan(pi/8)
  const double num = big ? mn - mx : mn;
  double den = big ? mn + mx : mx;
  den = den == 0.0 ? 1.0 : den;  // atan2(0, 0): t = 0
  const double t = fm_div(num, den);
  const double z = t * t;
  const double* A = FM_TAB(fm_atan_t);
  double p = A[0];
#pragma unroll
  for (int i = 1; i < 11; ++i) p = fm_fma(p, z, A[i]);
  double a = fm_fma(-t, z * p, t);
  // + pi/4 (hi + lo) when reduced
  a = big ? (K[10] + (a + K[11])) : a;
  a = ay > ax ? (K[12] - (a - K[13])) : a;
  a = fm_hi(x) < 0 ? (K[14] - (a - K[15])) : a;
  return fm_make(fm_hi(a) ^ (fm_hi(y) & 0x80000000), fm_lo(a));
}

}  // namespace fop
