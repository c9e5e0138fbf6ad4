// fp64 complex helpers for device code.  Principal branches match libstdc++'s std::complex
// (sqrt, log, pow(complex, real)), which is what the reference's characteristic functions use
// (SURVEY.md "Hard parts": replicate, don't improve).  No fast-math anywhere.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "fmath.cuh"

namespace fop {

// cx<T>: complex number over T = double or T = VD<N> (N independent values, see fmath.cuh).
// Mixed operations broadcast the scalar operand.
template <class T> struct cx { T x, y; };
using cd = cx<double>;
#define CX_HD __host__ __device__ __forceinline__

template <class T> CX_HD cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <class A, class B> CX_HD auto operator+(cx<A> a, cx<B> b) -> cx<decltype(a.x + b.x)> { return mk(a.x + b.x, a.y + b.y); }
template <class A, class B> CX_HD auto operator-(cx<A> a, cx<B> b) -> cx<decltype(a.x - b.x)> { return mk(a.x - b.x, a.y - b.y); }
template <class A> CX_HD cx<A> operator-(cx<A> a) { return mk(-a.x, -a.y); }
template <class A, class B> CX_HD auto operator*(cx<A> a, cx<B> b) -> cx<decltype(a.x * b.x)> {
  return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class A> CX_HD cx<A> operator*(double s, cx<A> a) { return mk(s * a.x, s * a.y); }
template <class A> CX_HD cx<A> operator*(cx<A> a, double s) { return mk(s * a.x, s * a.y); }
template <int N, class A> CX_HD cx<VD<N>> operator*(VD<N> s, cx<A> a) { return mk(s * a.x, s * a.y); }
template <int N, class A> CX_HD cx<VD<N>> operator*(cx<A> a, VD<N> s) { return mk(s * a.x, s * a.y); }
template <class A> CX_HD cx<A> operator+(cx<A> a, double s) { return mk(a.x + s, a.y); }
template <class A> CX_HD cx<A> operator+(double s, cx<A> a) { return mk(a.x + s, a.y); }
template <class A> CX_HD cx<A> operator-(cx<A> a, double s) { return mk(a.x - s, a.y); }
template <class A> CX_HD cx<A> operator-(double s, cx<A> a) { return mk(s - a.x, -a.y); }
template <int N> CX_HD cx<VD<N>> operator-(VD<N> s, cx<VD<N>> a) { return mk(s - a.x, -a.y); }
template <class A> CX_HD cx<A> conj(cx<A> a) { return mk(a.x, -a.y); }
template <class A> CX_HD A norm(cx<A> a) { return a.x * a.x + a.y * a.y; }

template <class A, class B> CX_HD auto cdiv(cx<A> a, cx<B> b) -> cx<decltype(a.x * b.x)> {
  const B inv = fm_rcp(b.x * b.x + b.y * b.y);
  return mk((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}
CX_HD cd cdiv(cd a, double s) { return mk(a.x / s, a.y / s); }

// The complex elementary functions below are straight-line code over fmath.cuh (no branches, no
// libdevice calls); with T = VD<N> the N evaluations advance operation by operation.
template <class T> CX_HD cx<T> cexp(cx<T> z) {
  const T e = fm_exp(z.x);
  T s, c;
  fm_sincos(z.y, &s, &c);
  return mk(e * c, e * s);
}
// v e^z (one multiplication fewer than v * cexp(z))
template <class T> CX_HD cx<T> cexp_scaled(cx<T> z, double v) {
  const T e = fm_exp(z.x) * v;
  T s, c;
  fm_sincos(z.y, &s, &c);
  return mk(e * c, e * s);
}
// e^{i t}
CX_HD cd cis(double t) {
  double s, c;
  fm_sincos(t, &s, &c);
  return mk(c, s);
}
// principal log: Im in (-pi, pi].  |z| is formed after scaling by a power of two so that tiny or
// huge CF values (FSTS theta takes log of phi ~ 1e-200) neither underflow nor overflow, matching
// libstdc++'s log(hypot(x, y)).  z = 0 gives -inf.
template <class T> CX_HD cx<T> clog(cx<T> z) {
  const T ax = fm_abs(z.x), ay = fm_abs(z.y);
  const T big = fm_sel(ax > ay, ax, ay);
  auto e = ((fm_hi(big) >> 20) & 0x7ff) - 1023;
  e = fm_sel(e < -1020, -1020, fm_sel(e > 1020, 1020, e));
  const T sc = fm_make((1023 - e) << 20, 0);  // 2^-e, exact
  const T xs = z.x * sc, ys = z.y * sc;
  const T n = fm_fma(xs, xs, ys * ys);
  const T lr = fm_fma(fm_i2d(e), FM_TAB(fm_k)[16], 0.5 * fm_log_pos(n));
  return mk(fm_sel(big == 0.0, -(double)INFINITY, lr), fm_atan2(z.y, z.x));
}
// q = a / w and lw = log w in one go for |w| = O(1) (the CIR clock's w = 1 - g (1 - e), which is 1
// at u = 0 and stays between ~1e-3 and ~10): |w|^2 is formed once, unscaled, and shared by the
// reciprocal and the logarithm; atan2 skips its power-of-two pre-scaling.
template <class T> CX_HD void cdiv_clog_unit(cx<T> a, cx<T> w, cx<T>* q, cx<T>* lw) {
  const T n = fm_fma(w.x, w.x, w.y * w.y);
  const T inv = fm_rcp(n);
  *q = mk((a.x * w.x + a.y * w.y) * inv, (a.y * w.x - a.x * w.y) * inv);
  *lw = mk(0.5 * fm_log_pos(n), fm_atan2_normal(w.y, w.x));
}
// principal square root (Re >= 0; on the negative real axis the sign of Im follows the sign of z.y)
__host__ __device__ __forceinline__ cd csqrt(cd z) {
  if (z.x == 0.0 && z.y == 0.0) return mk(0.0, z.y);
  const double r = sqrt(z.x * z.x + z.y * z.y);
  if (z.x >= 0.0) {
    const double t = sqrt(0.5 * (r + z.x));
    return mk(t, z.y / (2.0 * t));
  }
  const double t = sqrt(0.5 * (r - z.x));
  return mk(fabs(z.y) / (2.0 * t), copysign(t, z.y));
}
// the same principal square root without the built-in sqrt / division (two branch-free reciprocal square
// roots; 1/(2t) is the second one, so there is no division at all) -- the T-independent stage of the
// coefficient kernels (cf_prepare)
CX_HD cd csqrt_fast(cd z) {
  const double ax = fabs(z.x), ay = fabs(z.y);
  const double n = fma(z.x, z.x, z.y * z.y);
  const bool zero = !(n > 0.0) && n == 0.0;
  const double r = fm_sqrt_pos(zero ? 1.0 : n);       // |z|
  const double tt = 0.5 * (r + ax);                  // > 0
  double rt;
  const double t = fm_sqrt_pos(tt, &rt);
  const double o = (0.5 * ay) * rt;                  // |y| / (2 t)
  const bool pos = z.x >= 0.0;
  const double re = pos ? t : o, im = copysign(pos ? o : t, z.y);
  return mk(zero ? 0.0 : re, zero ? z.y : im);
}
// pow(complex, real): polar(exp(y log|z|), y arg z); 0^y = 0 (needed by the reference's own CGMY
// Carr-Madan test where (M - u) is exactly 0 at omega = 0, test.cpp:841 with alpha + 1 = 2.5 = M).
// For a positive real base this equals libstdc++'s real pow to rounding.
__host__ __device__ __forceinline__ cd cpow(cd z, double y) {
  const cd l = clog(z);
  const double m = fm_exp(y * l.x);
  double s, c;
  fm_sincos(y * l.y, &s, &c);
  const bool zero = z.x == 0.0 && z.y == 0.0;
  return mk(zero ? 0.0 : m * c, zero ? 0.0 : m * s);
}

}  // namespace fop
