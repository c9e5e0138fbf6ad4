// Drop-in for the hot function of the reference's OptionCalibration.h: optioncal::getObjFn_arr
// (:185-201).  The reference returns a closure objFn(theta...) -> double that a CPU optimiser calls
// ~25 x 1500 times; here the closure also accepts a whole population [P][np] and evaluates it in
// one kernel launch (fop_objfn_cf).
#ifndef FFTOP_OPTIONCALIBRATION_H
#define FFTOP_OPTIONCALIBRATION_H
#include <complex>
#include <vector>

#include "OptionPricing.h"

namespace optioncal {

constexpr double largeNumber = 5000;  // OptionCalibration.h:185

struct ObjFn {
  fop_model_t model;
  std::vector<std::complex<double>> phiHat;
  std::vector<double> uArray;
  double r, T;
  // one parameter vector -> scalar loss (the reference's call shape, test.cpp:1305-1306)
  double operator()(const std::vector<double>& theta) const { return population(theta, 1)[0]; }
  // P parameter vectors, row-major [P][np] -> P losses
  std::vector<double> population(const std::vector<double>& thetas, int P) const {
    auto& e = fftop::Engine::instance();
    std::vector<double> loss(P);
    e.check(fop_objfn_cf(e.get(), model, thetas.data(), P, uArray.data(),
                         reinterpret_cast<const double*>(phiHat.data()), (int)uArray.size(), r, T,
                         loss.data()));
    return loss;
  }
};

// getObjFn_arr(phiHat, cfFn, uArray): cfFn is a model family (the log CF of which is evaluated at
// 1 + i u_l with rate r and maturity T; generateCharts.cpp:231-250 uses r = 0)
inline ObjFn getObjFn_arr(std::vector<std::complex<double>>&& phiHat, const fftop::CF& family,
                          std::vector<double>&& uArray, double r = 0.0, double T = 1.0) {
  return ObjFn{family.model, std::move(phiHat), std::move(uArray), r, T};
}

}  // namespace optioncal
#endif
