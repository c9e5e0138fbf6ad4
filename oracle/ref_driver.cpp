// ORACLE / TEST INFRASTRUCTURE ONLY -- "reference" flavour.
//
// Compiles the reference's OWN headers, unchanged, from where they lie in /root/reference
// (OptionPricing.h, fft.h, fft.hpp, OptionCalibration.h, monotone_spline.h) against the shim
// headers in oracle/shims/ that restate the un-vendored third-party symbols, and exposes the
// result through the C ABI of oracle_abi.h.  Build recipe: oracle/Makefile (target `ref`);
// output only into oracle/_ref/ (git-ignored).  No reference source is copied into this repo.
//
// Batched entry points run `#pragma omp parallel for` over parameter sets with the reference
// call single-threaded inside (nested OpenMP regions are serialised) -- the most favourable
// arrangement for the CPU baseline.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "FunctionalUtilities.h"
#include "OptionPricing.h"        // /root/reference/OptionPricing.h
#include "fft.h"                  // /root/reference/fft.h (+ fft.hpp)
#include "CharacteristicFunctions.h"
#include "OptionCalibration.h"    // /root/reference/OptionCalibration.h

#include "oracle_abi.h"
#include "oracle_models.h"

using oramodel::Model;
using oramodel::cplx;

namespace {

// multi-step / American FSTS composed from the reference's fft/ifft
// (fft.hpp:49-61), getUDomain (OptionPricing.h:19-23) and the grid of FSTSGeneric (:165-167).
// One step is FSTSPrice's operator with T -> dt: v <- e^{-r dt} Re(ifft(fft(v) * phi_dt))/N.
std::vector<double> fstsMultiStep(const Model& mdlDt, int N, double xmax, double strike, double r,
                                  double dt, int steps, bool american, bool isPut) {
  const double dx = fangoost::computeDX(N, -xmax, xmax);
  const double vmax = M_PI / dx;
  const double du = 2.0 * vmax / N;
  std::vector<double> pay(N), v(N);
  for (int j = 0; j < N; ++j) {
    const double s = strike * exp(fangoost::getX(-xmax, dx, j));
    pay[j] = isPut ? (s < strike ? strike - s : 0.0) : (s > strike ? s - strike : 0.0);
  }
  std::vector<cplx> phi(N);
  for (int k = 0; k < N; ++k) phi[k] = mdlDt(cplx(0.0, optionprice::getUDomain(vmax, du, k)));
  const double disc = exp(-r * dt);
  v = pay;
  for (int s = 0; s < steps; ++s) {
    CArray x(N);
    for (int j = 0; j < N; ++j) x[j] = cplx(v[j], 0.0);
    x = fft(std::move(x));
    for (int k = 0; k < N; ++k) x[k] = x[k] * phi[k];
    x = ifft(std::move(x));
    for (int j = 0; j < N; ++j) {
      const double c = disc * x[j].real() / N;
      v[j] = american ? std::max(c, pay[j]) : c;
    }
  }
  return v;
}

std::vector<double> cosOne(const Model& mdl, const std::vector<double>& K, double S0, double r,
                           double T, int numU, int kind) {
  switch (kind) {
    case 0: return optionprice::FangOostCallPrice(S0, K, r, T, numU, Model(mdl));
    case 1: return optionprice::FangOostPutPrice(S0, K, r, T, numU, Model(mdl));
    case 2: return optionprice::FangOostCallDelta(S0, K, r, T, numU, Model(mdl));
    case 3: return optionprice::FangOostPutDelta(S0, K, r, T, numU, Model(mdl));
    case 4: return optionprice::FangOostCallGamma(S0, K, r, T, numU, Model(mdl));
    case 5: return optionprice::FangOostPutGamma(S0, K, r, T, numU, Model(mdl));
    case 6: return optionprice::FangOostCallTheta(S0, K, r, T, numU, Model(mdl));
    default: return optionprice::FangOostPutTheta(S0, K, r, T, numU, Model(mdl));
  }
}

}  // namespace

extern "C" {

const char* ora_flavor(void) { return "reference"; }
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
    std::vector<double> res =
        is_put ? optionprice::CarrMadanPut(N, ada, alpha, S0, r, T, Model(mdl))
               : optionprice::CarrMadanCall(N, ada, alpha, S0, r, T, Model(mdl));
    std::memcpy(out + (size_t)p * N, res.data(), sizeof(double) * N);
  }
  return 0;
}

int ora_cos(int base, int tc, const double* params, int P, const double* K_desc, int J,
            double S0, double r, const double* T, int M, int numU, int kind, double* out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !out || !K_desc || !T || J < 2 || kind < 0 || kind > 7) return 1;
  const std::vector<double> K(K_desc, K_desc + J);
  const long long rows = (long long)P * M;
#pragma omp parallel for schedule(dynamic)
  for (long long row = 0; row < rows; ++row) {
    const int p = (int)(row / M), m = (int)(row % M);
    Model mdl{base, tc, params + (size_t)p * np, r, T[m]};
    std::vector<double> res = cosOne(mdl, K, S0, r, T[m], numU, kind);
    std::memcpy(out + (size_t)row * J, res.data(), sizeof(double) * J);
  }
  return 0;
}

int ora_fsts(int base, int tc, const double* params, int P, int N, double xmax, double strike,
             double r, double T, int steps, int american, int payoff, int kind, double* out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !out || N < 2 || (N & (N - 1)) || steps < 1) return 1;
  if ((steps > 1 || american) && (tc != 0 || kind != 0)) return 4;
  const bool isPut = payoff == 1;
  auto getAssetPrice = [strike](const auto& logAsset) { return strike * exp(logAsset); };
  auto payoffFn = [strike, isPut](const auto& a) {
    return isPut ? (a < strike ? strike - a : 0.0) : (a > strike ? a - strike : 0.0);
  };
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p) {
    std::vector<double> res;
    if (steps == 1 && !american) {
      Model mdl{base, tc, params + (size_t)p * np, r, T};
      switch (kind) {
        case 0: res = optionprice::FSTSPrice(N, xmax, r, T, getAssetPrice, payoffFn, Model(mdl)); break;
        case 1: res = optionprice::FSTSDelta(N, xmax, r, T, getAssetPrice, payoffFn, Model(mdl)); break;
        case 2: res = optionprice::FSTSGamma(N, xmax, r, T, getAssetPrice, payoffFn, Model(mdl)); break;
        default: res = optionprice::FSTSTheta(N, xmax, r, T, getAssetPrice, payoffFn, Model(mdl)); break;
      }
    } else {
      const double dt = T / steps;
      Model mdl{base, tc, params + (size_t)p * np, r, dt};
      res = fstsMultiStep(mdl, N, xmax, strike, r, dt, steps, american != 0, isPut);
    }
    std::memcpy(out + (size_t)p * N, res.data(), sizeof(double) * N);
  }
  return 0;
}

int ora_objfn_cf(int base, int tc, const double* params, int P, const double* u,
                 const double* phiHat, int m, double r, double T, double* loss) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !loss || !u || !phiHat || m < 1) return 1;
  std::vector<cplx> ph(m);
  for (int l = 0; l < m; ++l) ph[l] = cplx(phiHat[2 * l], phiHat[2 * l + 1]);
  std::vector<double> uu(u, u + m);
  // getObjFn_arr (OptionCalibration.h:186-201): cfFn(theta) returns a log-CF closure
  auto objFn = optioncal::getObjFn_arr(
      std::vector<cplx>(ph),
      [=](const double* theta) {
        Model mdl{base, tc, theta, r, T};
        return [mdl](const cplx& cu) { return mdl.logCF(cu); };
      },
      std::vector<double>(uu));
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p) loss[p] = objFn(params + (size_t)p * np);
  return 0;
}

int ora_cos_loss(int base, int tc, const double* params, int P, const double* K_desc, int J,
                 double S0, double r, const double* T, int M, int numU, const double* mkt,
                 const double* w, double* loss, double* prices_out) {
  const int np = oramodel::numParams(base, tc);
  if (np == 0 || !params || !loss || !K_desc || !T || !mkt || J < 2) return 1;
  const std::vector<double> K(K_desc, K_desc + J);
#pragma omp parallel for schedule(dynamic)
  for (int p = 0; p < P; ++p) {
    double acc = 0.0;
    for (int m = 0; m < M; ++m) {
      Model mdl{base, tc, params + (size_t)p * np, r, T[m]};
      std::vector<double> res = optionprice::FangOostCallPrice(S0, K, r, T[m], numU, Model(mdl));
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
    CArray x(N);
    double* d = data + (size_t)b * 2 * N;
    for (int j = 0; j < N; ++j) x[j] = cplx(d[2 * j], d[2 * j + 1]);
    x = inverse ? ifft(std::move(x)) : fft(std::move(x));
    for (int j = 0; j < N; ++j) { d[2 * j] = x[j].real(); d[2 * j + 1] = x[j].imag(); }
  }
  return 0;
}

// the reference's own integrateFFT / integrateIFFT (fft.hpp:96-110); fn(currInput) returns the
// sample whose grid point currInput is (index recovered by rounding)
int ora_integrate_fft(const double* fn_samples, int batch, int N, double outputMin, double inputMin,
                      double inputMax, int inverse, double* out) {
  if (!fn_samples || !out || N < 2 || (N & (N - 1)) || !(inputMax > inputMin)) return 1;
  const double inputDx = (inputMax - inputMin) / (N - 1);
  for (int b = 0; b < batch; ++b) {
    const double* f = fn_samples + (size_t)b * 2 * N;
    auto fn = [&](double x) {
      long j = std::lround((x - inputMin) / inputDx);
      j = j < 0 ? 0 : (j > N - 1 ? N - 1 : j);
      return cplx(f[2 * j], f[2 * j + 1]);
    };
    // integrateIFFT's unused `Method` template parameter cannot be deduced (fft.hpp:104-107), so
    // the inverse case goes through genericIntegration with ifft, which is all integrateIFFT does
    const auto res = inverse ? genericIntegration(outputMin, inputMin, inputMax, fn, N, ifft)
                             : integrateFFT(outputMin, inputMin, inputMax, fn, N);
    double* o = out + (size_t)b * 2 * N;
    for (int j = 0; j < N; ++j) { o[2 * j] = res[j].real(); o[2 * j + 1] = res[j].imag(); }
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

int ora_fo_estimate(const double* strikes, const double* options, int n, double stock, double r,
                    double t, double minStrike, double maxStrike, int N, const double* u, int m,
                    double* phiHat_out) {
  const std::vector<double> k(strikes, strikes + n), o(options, options + n), uu(u, u + m);
  const auto est = optioncal::generateFOEstimate(k, o, stock, r, t, minStrike, maxStrike);
  const auto ph = est(N, uu);
  for (int l = 0; l < m; ++l) { phiHat_out[2 * l] = ph[l].real(); phiHat_out[2 * l + 1] = ph[l].imag(); }
  return 0;
}

int ora_carr_madan_strikes(int N, double ada, double S0, double* K_out) {
  std::vector<double> k = optionprice::getCarrMadanStrikes(ada, S0, N);
  std::memcpy(K_out, k.data(), sizeof(double) * N);
  return 0;
}
int ora_fang_oost_strikes(double xMin, double xMax, double S0, int numX, double* K_out) {
  std::vector<double> k = optionprice::getFangOostStrike(xMin, xMax, S0, numX);
  std::memcpy(K_out, k.data(), sizeof(double) * numX);
  return 0;
}
int ora_strike_underlying(double xMin, double xMax, double K, int numX, double* S_out) {
  std::vector<double> s = optionprice::getStrikeUnderlying(xMin, xMax, K, numX);
  std::memcpy(S_out, s.data(), sizeof(double) * numX);
  return 0;
}

}  // extern "C"
