// CPU accuracy test of csrc/fmath.cuh against long double libm (built by tests/test_fmath_host.py).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include "fmath.cuh"
using namespace fop;

static double ulp_of(long double v) {
  double d = (double)fabsl(v);
  if (d == 0.0) return 4.9e-324;
  return nextafter(d, INFINITY) - d;
}
static double ulps(double got, long double ref) { return (double)(fabsl((long double)got - ref) / ulp_of(ref)); }

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  int bad = 0;
  double m;
  // exp
  m = 0;
  for (int i = 0; i < 2000000; ++i) {
    double x = (i & 1) ? (U(rng) - 0.5) * 1400.0 : (U(rng) - 0.5) * 4.0;
    m = fmax(m, ulps(fm_exp(x), expl((long double)x)));
  }
  printf("exp     max ulp %.3f\n", m); if (!(m < 1.5)) ++bad;
  if (fm_exp(-800.0) != 0.0 || !std::isinf(fm_exp(800.0)) || !std::isnan(fm_exp(NAN))) { printf("exp specials\n"); ++bad; }
  // sincos
  double ms = 0, mc = 0;
  for (int i = 0; i < 2000000; ++i) {
    double sc = (i % 3 == 0) ? 1e9 : (i % 3 == 1) ? 1e5 : 10.0;
    double x = (U(rng) - 0.5) * 2 * sc, s, c;
    fm_sincos(x, &s, &c);
    ms = fmax(ms, ulps(s, sinl((long double)x)));
    mc = fmax(mc, ulps(c, cosl((long double)x)));
  }
  printf("sin     max ulp %.3f\ncos     max ulp %.3f\n", ms, mc); if (!(ms < 2.0 && mc < 2.0)) ++bad;
  // log
  m = 0;
  for (int i = 0; i < 2000000; ++i) {
    double x = (i & 1) ? exp((U(rng) - 0.5) * 1380.0) : 1.0 + (U(rng) - 0.5) * 0.9;
    m = fmax(m, ulps(fm_log_pos(x), logl((long double)x)));
  }
  printf("log     max ulp %.3f\n", m); if (!(m < 1.5)) ++bad;
  // atan2
  m = 0;
  for (int i = 0; i < 2000000; ++i) {
    double sx = exp((U(rng) - 0.5) * 40.0), sy = exp((U(rng) - 0.5) * 40.0);
    double x = (U(rng) - 0.5) * sx, y = (U(rng) - 0.5) * sy;
    m = fmax(m, ulps(fm_atan2(y, x), atan2l((long double)y, (long double)x)));
  }
  printf("atan2   max ulp %.3f\n", m); if (!(m < 2.5)) ++bad;
  for (int i = 0; i < 200000; ++i) {  // subnormal / huge operands
    double sx = (i & 1) ? 1e-310 : 1e300, x = (U(rng) - 0.5) * sx, y = (U(rng) - 0.5) * sx;
    m = fmax(m, ulps(fm_atan2(y, x), atan2l((long double)y, (long double)x)));
  }
  printf("atan2   max ulp %.3f (incl. subnormal/huge)\n", m); if (!(m < 2.5)) ++bad;
  if (fm_atan2(0.0, 1.0) != 0.0 || fabs(fm_atan2(0.0, -1.0) - M_PI) > 1e-15 || fm_atan2(0.0, 0.0) != 0.0 ||
      fabs(fm_atan2(1.0, 0.0) - M_PI / 2) > 1e-15 || fabs(fm_atan2(-1.0, 0.0) + M_PI / 2) > 1e-15) { printf("atan2 specials\n"); ++bad; }
  // div
  m = 0;
  for (int i = 0; i < 1000000; ++i) {
    double a = (U(rng) - 0.5) * exp((U(rng) - 0.5) * 200.0), b = (U(rng) + 1e-3) * exp((U(rng) - 0.5) * 200.0);
    m = fmax(m, ulps(fm_div(a, b), (long double)a / (long double)b));
  }
  printf("div     max ulp %.3f\n", m); if (!(m < 1.01)) ++bad;
  // sqrt / rsqrt (branch-free, coarse seed)
  m = 0;
  double mr = 0;
  for (int i = 0; i < 2000000; ++i) {
    double x = exp((U(rng) - 0.5) * 600.0);
    m = fmax(m, ulps(fm_sqrt_pos(x), sqrtl((long double)x)));
    mr = fmax(mr, ulps(fm_rsqrt_pos(x), 1.0L / sqrtl((long double)x)));
  }
  printf("sqrt    max ulp %.3f\nrsqrt   max ulp %.3f\n", m, mr); if (!(m < 1.01 && mr < 1.5)) ++bad;
  printf(bad ? "FAIL %d\n" : "PASS\n", bad);
  return bad != 0;
}
