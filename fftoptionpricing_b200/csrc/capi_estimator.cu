// C ABI: fop_fo_estimate -- market quotes -> empirical log-CF estimate phiHat (host only; replaces
// optioncal::generateFOEstimate, OptionCalibration.h:157-184, with getOptionSpline :105-156 and
// monotone_spline.h).  One-off per market snapshot, O(|u| N) on the host; its output feeds
// fop_objfn_cf.  The implementation is the header-only C++ in include/fftop/OptionCalibration.h.
#include <complex>
#include <vector>

#include "../../include/fftop/OptionCalibration.h"

extern "C" int fop_fo_estimate(const double* strikes, const double* options, int n, double stock,
                               double r, double t, double minStrike, double maxStrike, int N,
                               const double* u, int m, double* phiHat_out) {
  if (!strikes || !options || !u || !phiHat_out || n < 3 || N < 3 || m < 1) return FOP_INVALID_ARG;
  const std::vector<double> k(strikes, strikes + n), o(options, options + n), uu(u, u + m);
  const auto est = optioncal::generateFOEstimate(k, o, stock, r, t, minStrike, maxStrike);
  const std::vector<std::complex<double>> ph = est(N, uu);
  for (int l = 0; l < m; ++l) {
    phiHat_out[2 * l] = ph[l].real();
    phiHat_out[2 * l + 1] = ph[l].imag();
  }
  return FOP_OK;
}
