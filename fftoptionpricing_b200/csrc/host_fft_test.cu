// CPU unit test of the FFT index algebra in fft_smem.cuh (threads emulated sequentially, one pass
// at a time).  Not a product path: built and run only by tests/test_fft_host.py.
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>
#include "fft_smem.cuh"
using namespace fop;
typedef std::complex<double> C;
static const double PI = 3.14159265358979323846264338327950288;

static double run(int n, int sign, bool dit) {
  const int N = 1 << n;
  std::vector<double2> tw(N);
  for (int k = 0; k < N; ++k) { tw[k].x = cos(-2 * PI * k / N); tw[k].y = sin(-2 * PI * k / N); }
  std::vector<C> x(N), ref(N);
  unsigned s = 12345u + n;
  for (int j = 0; j < N; ++j) {
    s = s * 1664525u + 1013904223u; double a = (s >> 8) / 16777216.0 - .5;
    s = s * 1664525u + 1013904223u; double b = (s >> 8) / 16777216.0 - .5;
    x[j] = C(a, b);
  }
  // reference DFT: direct long double sum with a long double twiddle table
  std::vector<long double> wc(N), ws(N);
  for (int e = 0; e < N; ++e) {
    const long double ang = sign * 2.0L * 3.14159265358979323846264338327950288L * e / N;
    wc[e] = cosl(ang);
    ws[e] = sinl(ang);
  }
  for (int k = 0; k < N; ++k) {
    long double re = 0, im = 0;
    int e = 0;
    for (int j = 0; j < N; ++j) {
      re += x[j].real() * wc[e] - x[j].imag() * ws[e];
      im += x[j].real() * ws[e] + x[j].imag() * wc[e];
      e += k;
      if (e >= N) e -= N;
    }
    ref[k] = C((double)re, (double)im);
  }
  std::vector<cd> sm(fft_padded_size(N));
  const int np = fft_num_passes(n), T = N / 16;
  double err = 0, mx = 0;
  if (!dit) {
    for (int j = 0; j < N; ++j) sm[fft_pad(j)] = mk(x[j].real(), x[j].imag());
    for (int p = 0; p < np; ++p)
      for (int b = 0; b < T; ++b) {
        if (sign < 0) fft_dif_pass<-1>(sm.data(), b, n, p, tw.data());
        else fft_dif_pass<1>(sm.data(), b, n, p, tw.data());
      }
    for (int k = 0; k < N; ++k) {
      const cd v = sm[fft_pad(fft_pos(k, n))];
      err = fmax(err, std::abs(C(v.x, v.y) - ref[k]));
      mx = fmax(mx, std::abs(ref[k]));
      if (fft_inv_pos(fft_pos(k, n), n) != k) return 1e9;
    }
  } else {
    for (int j = 0; j < N; ++j) sm[fft_pad(fft_pos(j, n))] = mk(x[j].real(), x[j].imag());
    for (int p = 0; p < np; ++p)
      for (int b = 0; b < T; ++b) {
        if (sign < 0) fft_dit_pass<-1>(sm.data(), b, n, p, tw.data());
        else fft_dit_pass<1>(sm.data(), b, n, p, tw.data());
      }
    for (int k = 0; k < N; ++k) {
      const cd v = sm[fft_pad(k)];
      err = fmax(err, std::abs(C(v.x, v.y) - ref[k]));
      mx = fmax(mx, std::abs(ref[k]));
    }
  }
  return err / mx;
}

int main() {
  int bad = 0;
  for (int n = 4; n <= 13; ++n)
    for (int sign = -1; sign <= 1; sign += 2)
      for (int dit = 0; dit < 2; ++dit) {
        const double e = run(n, sign, dit != 0);
        printf("n=%2d sign=%+d %s rel_err=%.3e\n", n, sign, dit ? "DIT" : "DIF", e);
        if (!(e < 1e-14)) ++bad;
      }
  // register-level power tree
  {
    cd p[16];
    const double th = -2 * PI * 37 / 4096;
    tw_powers16(mk(cos(th), sin(th)), p);
    double e = 0;
    for (int k = 1; k < 16; ++k) e = fmax(e, hypot(p[k].x - cos(k * th), p[k].y - sin(k * th)));
    printf("tw_powers16 max err %.3e\n", e);
    if (!(e < 2e-15)) ++bad;
  }
  printf(bad ? "FAIL %d\n" : "PASS\n", bad);
  return bad != 0;
}
