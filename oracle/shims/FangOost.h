// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Shim for the un-vendored third-party header `FangOost.h`
// (phillyfan1138/FangOost @5919d52d5a59b317878b61115acd6cda3a91b97f, pinned at
// /root/reference/package.sh:23).  Source NOT in /root/reference.  Restated from the
// published COS method (Fang & Oosterlee 2008, eq. 22 / 30) and the reference's call sites:
//   computeDX / getX                 OptionPricing.h:86-88,94,100,122,165,172,182,605
//   computeExpectationVectorLevy     OptionPricing.h:336-345
//     cf   receives the COMPLEX argument i*u_k          (OptionPricing.h:338-340)
//     vK   receives the REAL u_k, the x value and k     (OptionPricing.h:341-343)
//     out  receives (value, x, index)                   (OptionPricing.h:377-379)
// Numerically pinned by the reference's own 12-digit constants (test.cpp:825, :1065).
#ifndef FFTOP_ORACLE_SHIM_FANGOOST_H
#define FFTOP_ORACLE_SHIM_FANGOOST_H
#include <complex>
#include <cmath>
#include <vector>
#include "FunctionalUtilities.h"
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace fangoost {

template <typename Index, typename Number>
auto computeDX(const Index& n, const Number& xMin, const Number& xMax) {
  return (xMax - xMin) / (static_cast<double>(n) - 1.0);
}
template <typename Number, typename Index>
auto getX(const Number& xMin, const Number& dx, const Index& index) {
  return xMin + dx * static_cast<double>(index);
}
template <typename Number>
auto computeDU(const Number& xMin, const Number& xMax) {
  return M_PI / (xMax - xMin);
}
template <typename Number>
auto computeCP(const Number& xMin, const Number& xMax) {
  return 2.0 / (xMax - xMin);
}

// val_j = sum'_k Re( cf(i u_k) cp e^{i u_k (x_j - xMin)} ) vK(u_k, x_j, k),  first term halved.
template <typename Array, typename Index, typename CF, typename VK, typename Out>
auto computeExpectationVectorLevy(Array&& xValues, const Index& numU, CF&& cf, VK&& vK,
                                  Out&& out) {
  const double xMin = xValues.front();
  const double xMax = xValues.back();
  const double du = computeDU(xMin, xMax);
  const double cp = computeCP(xMin, xMax);
  std::vector<std::complex<double>> dcf =
      futilities::for_each_parallel(0, static_cast<int>(numU), [&](const auto& k) {
        return cf(std::complex<double>(0.0, du * k)) * cp;
      });
  if (!dcf.empty()) dcf[0] *= 0.5;
  return futilities::for_each_parallel(
      std::move(xValues), [&](const auto& x, const auto& j) {
        const double val = futilities::sum(dcf, [&](const auto& d, const auto& k) {
          const double u = du * k;
          return (d * std::exp(std::complex<double>(0.0, u * (x - xMin)))).real() *
                 vK(u, x, k);
        });
        return out(val, x, j);
      });
}

}  // namespace fangoost
#endif
