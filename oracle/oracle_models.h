// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// Builds the characteristic-function closures the reference's tests build by hand
// (test.cpp:54,570,731,887-905,989,1050,1090,1119; generateCharts.cpp:190-380) from the
// (base, time_change, params) descriptor of include/fftop_b200.h.  The closures call the
// chfunctions:: shim (oracle/shims/CharacteristicFunctions.h) exactly the way the tests do.
#ifndef FFTOP_ORACLE_MODELS_H
#define FFTOP_ORACLE_MODELS_H
#include <complex>
#include <cmath>
#include "CharacteristicFunctions.h"

namespace oramodel {
typedef std::complex<double> cplx;

inline int numBase(int base) { return base == 0 ? 1 : base == 1 ? 4 : base == 2 ? 5 : base == 3 ? 3 : 0; }
inline int numParams(int base, int tc) {
  const int nb = numBase(base);
  if (nb == 0 || tc < 0 || tc > 2) return 0;
  return nb + (tc == 0 ? 0 : 4);
}

struct Model {
  int base, tc;
  const double* p;
  double r, T;

  // risk-neutral Levy exponent per unit time with rate `rate` and diffusion `sig`
  cplx baseLogRNCF(const cplx& u, double rate, double sig) const {
    switch (base) {
      case 0: return chfunctions::gaussLogCF(u, rate - sig * sig * .5, sig);       // test.cpp:570
      case 1: return chfunctions::mertonLogRNCF(u, p[1], p[2], p[3], rate, sig);   // test.cpp:1119
      case 2: return chfunctions::cgmyLogRNCF(u, p[1], p[2], p[3], p[4], rate, sig);  // :731
      default: return chfunctions::vgLogRNCF(u, p[0], p[1], p[2], rate);  // VG: no diffusion part
    }
  }
  // log of the characteristic function of log(S_T/S_0)
  cplx logCF(const cplx& u) const {
    const double sig = base == 3 ? 0.0 : p[0];  // Brownian volatility (variance gamma has none)
    if (tc == 0) return baseLogRNCF(u, r, sig) * T;
    const double* q = p + numBase(base);
    const double speed = q[0], adaV = q[1], rho = q[2], v0Hat = q[3];
    if (tc == 1)  // test.cpp:887-894, 1049-1058; generateCharts.cpp:231-250
      return r * T * u + chfunctions::cirLogMGF(-baseLogRNCF(u, 0.0, sig), speed,
                                                speed - adaV * rho * u * sig, adaV, T, v0Hat);
    // tc == 2: only the diffusion is time changed (test.cpp:896-905)
    cplx jumps(0.0, 0.0);
    if (base != 0) jumps = baseLogRNCF(u, 0.0, 0.0) * T;
    return r * T * u +
           chfunctions::cirLogMGF(-chfunctions::gaussLogCF(u, -sig * sig * .5, sig), speed,
                                  speed - adaV * rho * u * sig, adaV, T, v0Hat) +
           jumps;
  }
  cplx operator()(const cplx& u) const { return exp(logCF(u)); }
};
}  // namespace oramodel
#endif
