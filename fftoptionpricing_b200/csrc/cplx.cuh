// fp64 complex helpers for device code.  Principal branches match libstdc++'s std::complex
// (sqrt, log, pow(complex, real)), which is what the reference's characteristic functions use
// (SURVEY.md "Hard parts": replicate, don't improve).  No fast-math anywhere.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "fmath.cuh"

namespace fop {

struct cd {
  double x, y;
};

__host__ __device__ __forceinline__ cd mk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__host__ __device__ __forceinline__ cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cd operator-(cd a) { return mk(-a.x, -a.y); }
__host__ __device__ __forceinline__ cd operator*(cd a, cd b) {
  return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cd operator*(double s, cd a) { return mk(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cd operator*(cd a, double s) { return mk(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cd operator+(cd a, double s) { return mk(a.x + s, a.y); }
__host__ __device__ __forceinline__ cd operator+(double s, cd a) { return mk(a.x + s, a.y); }
__host__ __device__ __forceinline__ cd operator-(cd a, double s) { return mk(a.x - s, a.y); }
__host__ __device__ __forceinline__ cd operator-(double s, cd a) { return mk(s - a.x, -a.y); }
__host__ __device__ __forceinline__ cd conj(cd a) { return mk(a.x, -a.y); }
__host__ __device__ __forceinline__ double norm(cd a) { return a.x * a.x + a.y * a.y; }

__host__ __device__ __forceinline__ cd cdiv(cd a, cd b) {
  const double inv = fm_rcp(b.x * b.x + b.y * b.y);
  return mk((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}
__host__ __device__ __forceinline__ cd cdiv(cd a, double s) { return mk(a.x / s, a.y / s); }

// The complex elementary functions below are straight-line code over fmath.cuh (no branches, no
// libdevice calls), so several independent evaluations interleave in one basic block.
__host__ __device__ __forceinline__ cd cexp(cd z) {
  const double e = fm_exp(z.x);
  double s, c;
  fm_sincos(z.y, &s, &c);
  return mk(e * c, e * s);
}
// e^{i t}
__host__ __device__ __forceinline__ cd cis(double t) {
  double s, c;
  fm_sincos(t, &s, &c);
  return mk(c, s);
}
// principal log: Im in (-pi, pi].  |z| is formed after scaling by a power of two so that tiny or
// huge CF values (FSTS theta takes log of phi ~ 1e-200) neither underflow nor overflow, matching
// libstdc++'s log(hypot(x, y)).  z = 0 gives -inf.
__host__ __device__ __forceinline__ cd clog(cd z) {
  const double ax = fabs(z.x), ay = fabs(z.y);
  const double big = ax > ay ? ax : ay;
  int e = ((fm_hi(big) >> 20) & 0x7ff) - 1023;
  e = e < -1020 ? -1020 : (e > 1020 ? 1020 : e);
  const double sc = fm_make((1023 - e) << 20, 0);  // 2^-e, exact
  const double xs = z.x * sc, ys = z.y * sc;
  const double n = fm_fma(xs, xs, ys * ys);
  const double lr = fm_fma((double)e, FM_TAB(fm_k)[16], 0.5 * fm_log_pos(n));
  return mk(big == 0.0 ? -INFINITY : lr, fm_atan2(z.y, z.x));
}
// principal square root (Re >= 0; on the negative real axis the sign of Im follows the sign of z.y)
__device__ __forceinline__ cd csqrt(cd z) {
  if (z.x == 0.0 && z.y == 0.0) return mk(0.0, z.y);
  const double r = sqrt(z.x * z.x + z.y * z.y);
  if (z.x >= 0.0) {
    const double t = sqrt(0.5 * (r + z.x));
    return mk(t, z.y / (2.0 * t));
  }
  const double t = sqrt(0.5 * (r - z.x));
  return mk(fabs(z.y) / (2.0 * t), copysign(t, z.y));
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
