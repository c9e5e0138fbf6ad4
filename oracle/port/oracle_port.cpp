// ORACLE / TEST INFRASTRUCTURE ONLY -- "port" flavour.
//
// Independent CPU restatement of the reference's hot path; needs NO reference source, so it
// builds on the GPU box too.  It is pinned (tests/test_oracle.py) against
//   (1) the "reference" flavour (the reference's own headers compiled here, ref_driver.cpp), and
//   (2) the committed golden vectors in tests/golden/ that were generated from (1).
// Each function cites the reference file:line it restates (paths relative to /root/reference).
// The characteristic functions come from oracle/shims/CharacteristicFunctions.h via
// oracle_models.h (restatement of the un-vendored CharacteristicFunctions library).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#include "oracle_abi.h"
#include "oracle_models.h"

using oramodel::Model;
using oramodel::cplx;

namespace {

// ---- fft.hpp:5-46 -- in-place radix-2 decimation in frequency, twiddle by running product,
// then bit-reversal.  `w` is e^{-+ i pi/N}; it is squared once per stage (fft.hpp:13) and the
// per-butterfly twiddle is accumulated by repeated multiplication (fft.hpp:24) -- the rounding
// behaviour of that recurrence is part of what parity is measured against, so it is kept.
void radix2_dif(std::vector<cplx>& x, cplx w) {
  const unsigned n = (unsigned)x.size();
  for (unsigned half = n >> 1, span = n; half >= 1; span = half, half >>= 1) {
    w = w * w;
    cplx tw(1.0, 0.0);
    for (unsigned l = 0; l < half; ++l) {
      for (unsigned a = l; a < n; a += span) {
        const unsigned b = a + half;
        const cplx diff = x[a] - x[b];
        x[a] = x[a] + x[b];
        x[b] = diff * tw;
      }
      tw *= w;
    }
  }
  unsigned bits = 0;
  while ((1u << bits) < n) ++bits;
  for (unsigned a = 0; a < n; ++a) {
    unsigned b = 0;
    for (unsigned t = 0; t < bits; ++t) b |= ((a >> t) & 1u) << (bits - 1 - t);
    if (b > a) std::swap(x[a], x[b]);
  }
}
// fft.hpp:56-61 (forward, e^{-i pi/N}) and fft.hpp:49-54 (inverse, e^{+i pi/N}); no 1/N.
void fwd(std::vector<cplx>& x) {
  const double th = M_PI / (double)x.size();
  radix2_dif(x, cplx(cos(th), sin(-th)));
}
void inv(std::vector<cplx>& x) {
  const double th = M_PI / (double)x.size();
  radix2_dif(x, cplx(cos(th), sin(th)));
}

// ---- OptionPricing.h:582-607 (CarrMadanGeneric) + :564-567 (CallAug) + :620-656 wrappers
void carrMadanOne(const Model& cf, int N, double ada, double alpha, double S0, double r, double T,
                  bool isPut, double* out) {
  const double discount = exp(-r * T);
  const double b = M_PI / ada;        // getMaxK  :31-33
  const double lambda = 2.0 * b / N;  // getLambda :40-42
  std::vector<cplx> x(N);
  for (int j = 0; j < N; ++j) {
    const cplx v(0.0, j * ada);
    const double pm = (j % 2 == 0) ? -1.0 : 1.0;
    const cplx aug = cf(v + (alpha + 1.0)) / (alpha * alpha + alpha + v * v + (2 * alpha + 1) * v);
    x[j] = discount * aug * exp(cplx(0.0, b * j * ada)) * (3.0 + pm);  // :599
  }
  x[0] = x[0] / 2.0;  // :594
  fwd(x);
  for (int i = 0; i < N; ++i) {
    const double k = -b + lambda * i;
    double val = S0 * x[i].real() * exp(-alpha * k) * ada / (M_PI * 3.0);  // :605
    if (isPut) val = val - S0 + S0 * exp(k) * discount;                    // :652, :121-123
    out[i] = val;
  }
}

// ---- OptionPricing.h:284-300 (chiK, phiK) with c = a = xMin, d = 0
double chiK(double a, double c, double d, double u) {
  const double ed = exp(d), ec = exp(c);
  return (cos(u * (d - a)) * ed - cos(u * (c - a)) * ec + u * sin(u * (d - a)) * ed -
          u * sin(u * (c - a)) * ec) / (1.0 + u * u);
}
double phiK(double a, double c, double d, double u, int k) {
  return k == 0 ? d - c : (sin(u * (d - a)) - sin(u * (c - a))) / u;
}

// ---- OptionPricing.h:44-73 transforms, :325-346 FangOostGeneric, :358-559 entry points, and the
// COS sum of the absent fangoost::computeExpectationVectorLevy (Fang & Oosterlee 2008 eq. 30).
void cosOne(const Model& cf, const double* K, int J, double S0, double r, double T, int numU,
            int kind, double* out) {
  const double discount = exp(-r * T);
  std::vector<double> x(J);
  for (int j = 0; j < J; ++j) x[j] = log(S0 / K[j]);  // getXFromK :110-117
  const double xMin = x.front(), xMax = x.back();
  const double du = M_PI / (xMax - xMin), cp = 2.0 / (xMax - xMin);
  const int g = kind >> 1;  // 0 price, 1 delta, 2 gamma, 3 theta
  std::vector<cplx> d(numU);
  std::vector<double> vk(numU);
  for (int k = 0; k < numU; ++k) {
    const cplx u(0.0, du * k);
    cplx c = cf(u);
    if (g == 1) c = c * u;                                                  // :52-57
    else if (g == 2) c = -c * (u * (1.0 - u));                              // :60-65
    else if (g == 3) c = c.real() > 0.0 ? -(log(c) - r) * c : cplx(0.0);    // :68-73
    d[k] = c * cp;
    vk[k] = phiK(xMin, xMin, 0.0, du * k, k) - chiK(xMin, xMin, 0.0, du * k);  // :342
  }
  d[0] *= 0.5;
  for (int j = 0; j < J; ++j) {
    double val = 0.0;
    for (int k = 0; k < numU; ++k)
      val += (d[k] * exp(cplx(0.0, du * k * (x[j] - xMin)))).real() * vk[k];
    double o;
    switch (kind) {
      case 0: o = (val - 1.0) * discount * K[j] + S0; break;    // :378
      case 1: o = val * discount * K[j]; break;                 // :414
      case 2: o = val * discount * K[j] / S0 + 1.0; break;      // :439
      case 3: o = val * discount * K[j] / S0; break;            // :464
      case 4: case 5: o = val * discount * K[j] / (S0 * S0); break;  // :541, :558
      case 6: o = (val - r) * discount * K[j]; break;           // :490
      default: o = val * discount * K[j]; break;                // :515
    }
    out[j] = o;
  }
}

// ---- OptionPricing.h:155-184 FSTSGeneric, :185-281 entry points, :19-23 getUDomain,
// and SURVEY.md A.4 for steps > 1 / american.
void fstsOne(int base, int tc, const double* theta, int N, double xmax, double strike, double r,
             double T, int steps, bool american, bool isPut, int kind, double* out) {
  const double dx = (2.0 * xmax) / (N - 1.0);  // computeDX(N,-xmax,xmax) :165
  const double vmax = M_PI / dx;               // :166
  const double du = 2.0 * vmax / N;            // :167
  const double dt = T / steps;
  const Model cf{base, tc, theta, r, (steps == 1 && !american) ? T : dt};
  std::vector<double> asset(N), pay(N);
  for (int j = 0; j < N; ++j) {
    asset[j] = strike * exp(-xmax + dx * j);
    pay[j] = isPut ? (asset[j] < strike ? strike - asset[j] : 0.0)
                   : (asset[j] > strike ? asset[j] - strike : 0.0);
  }
  std::vector<cplx> mult(N);
  for (int k = 0; k < N; ++k) {
    double w = du * k;
    if (w > vmax) w -= 2.0 * vmax;  // getUDomain :20-23 (strict >)
    const cplx u(0.0, w);
    cplx c = cf(u);
    if (kind == 1) c = c * u;
    else if (kind == 2) c = -c * (u * (1.0 - u));
    else if (kind == 3) c = c.real() > 0.0 ? -(log(c) - r) * c : cplx(0.0);
    mult[k] = c;
  }
  std::vector<double> v(pay);
  const double disc = exp(-r * ((steps == 1 && !american) ? T : dt));
  for (int s = 0; s < steps; ++s) {
    std::vector<cplx> x(N);
    for (int j = 0; j < N; ++j) x[j] = cplx(v[j], 0.0);
    fwd(x);
    for (int k = 0; k < N; ++k) x[k] = x[k] * mult[k];
    inv(x);
    for (int j = 0; j < N; ++j) {
      const double c = disc * x[j].real() / N;  // :202
      v[j] = american ? std::max(c, pay[j]) : c;
    }
  }
  for (int j = 0; j < N; ++j) {
    double o = v[j];
    if (kind == 1) o = o / asset[j];                      // :226
    else if (kind == 2) o = o / (asset[j] * asset[j]);    // :275
    out[j] = o;
  }
}

}  // namespace

extern "C" {

const char* ora_flavor(void) { return "port"; }
int ora_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ora_carr_madan(int base, int tc, const double* params, int P, int N, double ada,
                   double alpha, double S0, double r, double T, int is_put, double* out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !out || N < 2 || (N & (N - 1))) return 1;
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p) {
    Model mdl{base, tc, params + (size_t)p * np, r, T};
    carrMadanOne(mdl, N, ada, alpha, S0, r, T, is_put != 0, out + (size_t)p * N);
  }
  return 0;
}

int ora_cos(int base, int tc, const double* params, int P, const double* K_desc, int J,
            double S0, double r, const double* T, int M, int numU, int kind, double* out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !out || !K_desc || !T || J < 2 || kind < 0 || kind > 7) return 1;
  const long long rows = (long long)P * M;
#pragma omp parallel for schedule(dynamic)
  for (long long row = 0; row < rows; ++row) {
    const int p = (int)(row / M), m = (int)(row % M);
    Model mdl{base, tc, params + (size_t)p * np, r, T[m]};
    cosOne(mdl, K_desc, J, S0, r, T[m], numU, kind, out + (size_t)row * J);
  }
  return 0;
}

int ora_fsts(int base, int tc, const double* params, int P, int N, double xmax, double strike,
             double r, double T, int steps, int american, int payoff, int kind, double* out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !out || N < 2 || (N & (N - 1)) || steps < 1) return 1;
  if ((steps > 1 || american) && (tc != 0 || kind != 0)) return 4;
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p)
    fstsOne(base, tc, params + (size_t)p * np, N, xmax, strike, r, T, steps, american != 0,
            payoff == 1, kind, out + (size_t)p * N);
  return 0;
}

// OptionCalibration.h:185-201 getObjFn_arr
int ora_objfn_cf(int base, int tc, const double* params, int P, const double* u,
                 const double* phiHat, int m, double r, double T, double* loss) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !loss || !u || !phiHat || m < 1) return 1;
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p) {
    Model mdl{base, tc, params + (size_t)p * np, r, T};
    double acc = 0.0;
    for (int l = 0; l < m; ++l) {
      const cplx res = mdl.logCF(cplx(1.0, u[l]));
      acc += (std::isnan(res.real()) || std::isnan(res.imag()))
                 ? 5000.0
                 : std::norm(cplx(phiHat[2 * l], phiHat[2 * l + 1]) - res);
    }
    loss[p] = acc / (double)m;
  }
  return 0;
}

int ora_cos_loss(int base, int tc, const double* params, int P, const double* K_desc, int J,
                 double S0, double r, const double* T, int M, int numU, const double* mkt,
                 const double* w, double* loss, double* prices_out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !loss || !K_desc || !T || !mkt || J < 2) return 1;
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p) {
    std::vector<double> res(J);
    double acc = 0.0;
    for (int m = 0; m < M; ++m) {
      Model mdl{base, tc, params + (size_t)p * np, r, T[m]};
      cosOne(mdl, K_desc, J, S0, r, T[m], numU, 0, res.data());
      for (int j = 0; j < J; ++j) {
        const double d = res[j] - mkt[(size_t)m * J + j];
        acc += (w ? w[(size_t)m * J + j] : 1.0) * d * d;
      }
      if (prices_out)
        std::memcpy(prices_out + ((size_t)p * M + m) * J, res.data(), sizeof(double) * J);
    }
    loss[p] = acc / ((double)M * (double)J);
  }
  return 0;
}

int ora_fft(double* data, int batch, int N, int inverse) {
  if (!data || N < 2 || (N & (N - 1))) return 1;
  for (int b = 0; b < batch; ++b) {
    std::vector<cplx> x(N);
    double* d = data + (size_t)b * 2 * N;
    for (int j = 0; j < N; ++j) x[j] = cplx(d[2 * j], d[2 * j + 1]);
    if (inverse) inv(x); else fwd(x);
    for (int j = 0; j < N; ++j) { d[2 * j] = x[j].real(); d[2 * j + 1] = x[j].imag(); }
  }
  return 0;
}

// restatement of genericIntegration (fft.hpp:64-82)
int ora_integrate_fft(const double* fn_samples, int batch, int N, double outputMin, double inputMin,
                      double inputMax, int inverse, double* out) {
  if (!fn_samples || !out || N < 2 || (N & (N - 1)) || !(inputMax > inputMin)) return 1;
  const double range = inputMax - inputMin, inputDx = range / (N - 1);
  const double outputMax = outputMin + (N - 1) * 2.0 * M_PI / range, du = (outputMax - outputMin) / N;
  const cplx I(0.0, 1.0);
  for (int b = 0; b < batch; ++b) {
    const double* f = fn_samples + (size_t)b * 2 * N;
    std::vector<cplx> x(N);
    for (int j = 0; j < N; ++j) {
      const cplx v = cplx(f[2 * j], f[2 * j + 1]) * inputDx * std::exp(outputMin * inputDx * j * I) / 3.0;
      x[j] = (j == 0 || j == N - 1) ? v : (j % 2 == 0 ? 2.0 * v : 4.0 * v);
    }
    if (inverse) inv(x); else fwd(x);
    double* o = out + (size_t)b * 2 * N;
    for (int j = 0; j < N; ++j) {
      const double u = outputMin + j * du;
      const cplx r = x[j] * std::exp(u * inputMin * I);
      o[2 * j] = r.real(); o[2 * j + 1] = r.imag();
    }
  }
  return 0;
}

int ora_cf(int base, int tc, const double* params, double r, double T, double u_re, double u_im,
           double* cf_out2, double* logcf_out2) {
  if (oramodel::numParams(base, tc) == 0 || !params) return 1;
  Model mdl{base, tc, params, r, T};
  const cplx l = mdl.logCF(cplx(u_re, u_im));
  const cplx c = exp(l);
  if (cf_out2) { cf_out2[0] = c.real(); cf_out2[1] = c.imag(); }
  if (logcf_out2) { logcf_out2[0] = l.real(); logcf_out2[1] = l.imag(); }
  return 0;
}

int ora_fo_estimate(const double*, const double*, int, double, double, double, double, double, int,
                    const double*, int, double*) { return 4; }

// OptionPricing.h:120-138
int ora_carr_madan_strikes(int N, double ada, double S0, double* K_out) {
  const double b = M_PI / ada, lambda = 2.0 * b / N;
  for (int i = 0; i < N; ++i) K_out[i] = S0 * exp(-b + lambda * i);
  return 0;
}
// OptionPricing.h:92-105
int ora_fang_oost_strikes(double xMin, double xMax, double S0, int numX, double* K_out) {
  const double dx = (xMax - xMin) / (numX - 1.0);
  for (int i = 0; i < numX; ++i) K_out[i] = S0 * exp(-(xMin + dx * i));
  return 0;
}
// OptionPricing.h:84-90
int ora_strike_underlying(double xMin, double xMax, double K, int numX, double* S_out) {
  const double dx = (xMax - xMin) / (numX - 1.0);
  for (int i = 0; i < numX; ++i) S_out[i] = K * exp(xMin + dx * i);
  return 0;
}

}  // extern "C"
