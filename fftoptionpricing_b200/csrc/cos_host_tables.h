// Host-side tables of the COS path that depend only on (strikes, S0, maturities, numU): the payoff
// coefficients V_k and the common maturity grid.  Shared by capi_cos.cu and the CPU emulation of
// the device code in tests/hostemu.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace fop {

// chiK / phiK of OptionPricing.h:284-300 with c = a = xMin, d = 0  (V_k = phi_k - chi_k, :342)
inline double vk_put(double xMin, double u, int k) {
  const double a = xMin, c = xMin, d = 0.0;
  const double ed = std::exp(d), ec = std::exp(c);
  const double chi = (std::cos(u * (d - a)) * ed - std::cos(u * (c - a)) * ec +
                      u * std::sin(u * (d - a)) * ed - u * std::sin(u * (c - a)) * ec) /
                     (1.0 + u * u);
  const double phi = k == 0 ? d - c : (std::sin(u * (d - a)) - std::sin(u * (c - a))) / u;
  return phi - chi;
}

// Common grid of the maturities: T_m = n_m dt with integers 1 <= n_m < 4096, dt = min(T) / q, to
// within a few ulps (anything looser would change the prices).  Market maturities (weeks, months)
// are of this form; when they are not, n stays empty and every maturity gets its own exp.
inline bool maturity_grid(const std::vector<double>& T, double* dt_out, std::vector<int>* n) {
  const double tmin = *std::min_element(T.begin(), T.end());
  if (!(tmin > 0.0) || T.size() < 2) return false;
  for (int q = 1; q <= 64; ++q) {
    const double dt = tmin / q;
    bool ok = true;
    n->clear();
    for (double t : T) {
      const double ratio = t / dt, k = std::nearbyint(ratio);
      if (k < 1.0 || k >= 4096.0 || std::fabs(ratio - k) > 2e-15 * k) { ok = false; break; }
      n->push_back((int)k);
    }
    if (ok) { *dt_out = dt; return true; }
  }
  n->clear();
  return false;
}

// steps along the maturity list for pow_step (cos_coeff.cuh): d_m = n_m - n_{m-1} (n_{-1} = 0);
// empty when the list is not ascending (the grid shortcut is then not used)
inline std::vector<int> maturity_steps(const std::vector<int>& n) {
  std::vector<int> steps(n.size());
  int prev = 0;
  for (size_t m = 0; m < n.size(); prev = n[m], ++m) {
    if (n[m] < prev) return {};
    steps[m] = n[m] - prev;
  }
  return steps;
}

}  // namespace fop
