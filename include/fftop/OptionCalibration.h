// Drop-in for the hot function of the reference's OptionCalibration.h: optioncal::getObjFn_arr
// (:185-201).  The reference returns a closure objFn(theta...) -> double that a CPU optimiser calls
// ~25 x 1500 times; here the closure also accepts a whole population [P][np] and evaluates it in
// one kernel launch (fop_objfn_cf).
#ifndef FFTOP_OPTIONCALIBRATION_H
#define FFTOP_OPTIONCALIBRATION_H
#include <cmath>
#include <complex>
#include <tuple>
#include <utility>
#include <vector>

#include "OptionPricing.h"

// ---- host-side preprocessing: market quotes -> empirical log-CF estimate phiHat -----------------
//   These run once per market snapshot on the host -- O(|u| N) with N = 1024..4096
// -- and feed fop_objfn_cf.  Behaviour follows the reference's monotone_spline.h:9-79 and
// OptionCalibration.h:26-184, restated here as small classes instead of nested closures.
namespace spline {

// Monotonicity-preserving piecewise cubic through (xs, ys), xs ascending.  Tangents are the
// harmonic-type mean of the neighbouring secant slopes, zero where the secants change sign.
class MonotoneCubic {
 public:
  MonotoneCubic() = default;
  MonotoneCubic(std::vector<double> xs, std::vector<double> ys) : x_(std::move(xs)), y_(std::move(ys)) {
    const int n = (int)x_.size();
    std::vector<double> h(n - 1), sec(n - 1);
    for (int i = 0; i + 1 < n; ++i) {
      h[i] = x_[i + 1] - x_[i];
      sec[i] = (y_[i + 1] - y_[i]) / h[i];
    }
    t_.resize(n);
    t_[0] = sec[0];
    for (int i = 0; i + 2 < n; ++i) {
      if (sec[i] * sec[i + 1] <= 0) {
        t_[i + 1] = 0;
      } else {
        const double both = h[i] + h[i + 1];
        t_[i + 1] = 3 * both / ((both + h[i + 1]) / sec[i] + (both + h[i]) / sec[i + 1]);
      }
    }
    t_[n - 1] = sec[n - 2];
    q_.resize(n - 1);
    c_.resize(n - 1);
    for (int i = 0; i + 1 < n; ++i) {
      const double inv = 1.0 / h[i];
      const double bend = t_[i] + t_[i + 1] - 2.0 * sec[i];
      q_[i] = (sec[i] - t_[i] - bend) * inv;
      c_[i] = bend * inv * inv;
    }
  }
  double operator()(double x) const {
    const int n = (int)x_.size();
    if (x == x_.back()) return y_.back();
    // bisection over the knots 0 .. n-2; an exact knot hit returns its ordinate
    int lo = 0, hi = n - 2;
    while (lo <= hi) {
      const int mid = (int)std::floor(.5 * (lo + hi));
      if (x_[mid] < x) lo = mid + 1;
      else if (x_[mid] > x) hi = mid - 1;
      else return y_[mid];
    }
    const int seg = hi > 0 ? hi : 0;
    const double d = x - x_[seg], d2 = d * d;
    return y_[seg] + t_[seg] * d + q_[seg] * d2 + c_[seg] * d * d2;
  }

 private:
  std::vector<double> x_, y_, t_, q_, c_;
};
inline MonotoneCubic generateSpline(std::vector<double> xs, std::vector<double> ys) {
  return MonotoneCubic(std::move(xs), std::move(ys));
}

}  // namespace spline

namespace optioncal {

constexpr double largeNumber = 5000;  // OptionCalibration.h:185

// values / asset with `minV/asset` prepended and `maxV/asset` appended (OptionCalibration.h:62-78)
inline std::vector<double> transformPrices(const std::vector<double>& arr, double asset, double minV,
                                           double maxV) {
  std::vector<double> out;
  out.reserve(arr.size() + 2);
  out.push_back(minV / asset);
  for (double v : arr) out.push_back(v / asset);
  out.push_back(maxV / asset);
  return out;
}
// split `arr` by predicate(value, index) into (mapped-if-true, mapped-if-false)  (:80-93)
template <typename Pred, typename F1, typename F2>
std::tuple<std::vector<double>, std::vector<double>> filter(const std::vector<double>& arr,
                                                           const Pred& pred, F1&& yes, F2&& no) {
  std::vector<double> a, b;
  for (int i = 0; i < (int)arr.size(); ++i) {
    if (pred(arr[i], i)) a.emplace_back(yes(arr[i], i));
    else b.emplace_back(no(arr[i], i));
  }
  return std::make_tuple(a, b);
}
// index of the last normalised strike that is <= 1  (:95-103)
inline int getKThatIsBelowOne(const std::vector<double>& paddedStrikes) {
  const int n = (int)paddedStrikes.size();
  int i = 0;
  while (i < n && paddedStrikes[i] <= 1) ++i;
  return i - 1;
}

// Two-piece interpolant of normalised call prices in the normalised strike (:105-156): below the
// threshold a spline through price - intrinsic value, above it a spline through log(price).
class OptionSpline {
 public:
  OptionSpline(const std::vector<double>& strikes, const std::vector<double>& options, double stock,
               double discount, double minStrike, double maxStrike)
      : discount_(discount) {
    const std::vector<double> k = transformPrices(strikes, stock, minStrike, maxStrike);
    const std::vector<double> c = transformPrices(options, stock, stock - minStrike * discount, 0.0000001);
    const int at = getKThatIsBelowOne(k);
    threshold_ = (k[at] + k[at - 1]) * .5;
    std::vector<double> klo, khi, clo, chi;
    for (int i = 0; i < (int)k.size(); ++i) {
      if (k[i] < threshold_) {
        klo.push_back(k[i]);
        clo.push_back(c[i] - intrinsic(k[i]));
      } else {
        khi.push_back(k[i]);
        chi.push_back(std::log(c[i]));
      }
    }
    low_ = spline::MonotoneCubic(std::move(klo), std::move(clo));
    high_ = spline::MonotoneCubic(std::move(khi), std::move(chi));
  }
  double operator()(double k) const {
    return k < threshold_ ? low_(k) : std::exp(high_(k)) - intrinsic(k);
  }

 private:
  double intrinsic(double k) const {
    const double v = 1 - k * discount_;
    return v > 0.0 ? v : 0.0;
  }
  double discount_, threshold_;
  spline::MonotoneCubic low_, high_;
};
inline OptionSpline getOptionSpline(const std::vector<double>& strikes, const std::vector<double>& options,
                                    double stock, double discount, double minStrike, double maxStrike) {
  return OptionSpline(strikes, options, stock, discount, minStrike, maxStrike);
}

// phiHat(u) = log(1 + iu(1+iu) * int e^{iux} max(O(e^x/discount), 0) dx), Simpson over N nodes on
// [log(discount minK/S), log(discount maxK/S)]  (:42-55 DFT, :157-184 generateFOEstimate)
class FOEstimate {
 public:
  FOEstimate(const std::vector<double>& strikes, const std::vector<double>& options, double stock,
             double r, double t, double minStrike, double maxStrike)
      : discount_(std::exp(-r * t)),
        spline_(strikes, options, stock, discount_, minStrike, maxStrike),
        kmin_(minStrike / stock), kmax_(maxStrike / stock) {}
  std::vector<std::complex<double>> operator()(int N, const std::vector<double>& uArray) const {
    const std::complex<double> I(0, 1);
    const double xMin = std::log(discount_ * kmin_), xMax = std::log(discount_ * kmax_);
    const double dx = (xMax - xMin) / (N - 1);
    std::vector<double> f(N);
    for (int j = 0; j < N; ++j) {
      const double v = spline_(std::exp(xMin + dx * j) / discount_);
      const double wgt = (j == 0 || j == N - 1) ? 1.0 : (j % 2 == 0 ? 2.0 : 4.0);
      f[j] = (v > 0.0 ? v : 0.0) * wgt * dx / 3.0;
    }
    std::vector<std::complex<double>> out(uArray.size());
    for (size_t l = 0; l < uArray.size(); ++l) {
      const double u = uArray[l];
      std::complex<double> acc(0, 0);
      for (int j = 0; j < N; ++j) acc += std::exp(I * u * (xMin + dx * j)) * f[j];
      out[l] = std::log(1.0 + (u * I * (1.0 + u * I)) * acc);
    }
    return out;
  }

 private:
  double discount_;
  OptionSpline spline_;
  double kmin_, kmax_;
};
inline FOEstimate generateFOEstimate(const std::vector<double>& strikes, const std::vector<double>& options,
                                     double stock, double r, double t, double minStrike, double maxStrike) {
  return FOEstimate(strikes, options, stock, r, t, minStrike, maxStrike);
}

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
