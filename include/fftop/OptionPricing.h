// Drop-in C++ host header for the B200 engine: keeps the reference's entry points
// (namespace optionprice, reference OptionPricing.h) over the C ABI of fftop_b200.h.
//
// Differences from the reference, by necessity of running on a GPU:
//  * the fast path takes the characteristic function as a model descriptor fftop::CF (closed set:
//    exactly the CFs of test.cpp / generateCharts.cpp), evaluated on the device;
//  * the reference's own signatures -- ANY callable as characteristic function, payoff and asset map
//    (FangOostGeneric OptionPricing.h:325-333, CarrMadanGeneric :582-590, FSTSGeneric :155-163) --
//    are kept as overloads: the callable is evaluated by the host on the grid the library defines
//    and the samples go through fop_*_samples, everything else runs on the GPU.
// Everything else -- names, argument order and meaning, return by value as std::vector<double>
// (or the caller's strike container for COS, test.cpp:574 uses std::deque) -- is the reference's.
// Like the reference, no argument validation is reported to the caller except through
// fftop::Error exceptions for structural failures (no device, out of memory, unsupported size).
#ifndef FFTOP_OPTIONPRICING_H
#define FFTOP_OPTIONPRICING_H
#include <cmath>
#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../fftop_b200.h"
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace fftop {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

// one handle per host thread (handles are not thread-safe)
class Engine {
 public:
  explicit Engine(int device = 0) {
    const int st = fop_create(&h_, device);
    if (st != FOP_OK) throw Error(st, "fop_create failed: no usable CUDA device (no CPU fallback)");
  }
  ~Engine() { if (h_) fop_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  fop_handle_t get() const { return h_; }
  void check(int st) const { if (st != FOP_OK) throw Error(st, fop_last_error(h_)); }
  static Engine& instance() {
    static thread_local std::unique_ptr<Engine> e;
    if (!e) e.reset(new Engine(0));
    return *e;
  }
 private:
  fop_handle_t h_ = nullptr;
};

// model descriptor + one parameter set (layout: include/fftop_b200.h)
struct CF {
  fop_model_t model;
  std::vector<double> params;
};
inline CF blackScholes(double sigma) { return {{FOP_GAUSS, FOP_TC_NONE}, {sigma}}; }
inline CF merton(double sigma, double lambda, double muJ, double sigJ) {
  return {{FOP_MERTON, FOP_TC_NONE}, {sigma, lambda, muJ, sigJ}};
}
inline CF cgmy(double sigma, double C, double G, double M, double Y) {
  return {{FOP_CGMY, FOP_TC_NONE}, {sigma, C, G, M, Y}};
}
// variance gamma (not in the reference: include/fftop_b200.h, FOP_VG)
inline CF varianceGamma(double sigma, double theta, double nu) { return {{FOP_VG, FOP_TC_NONE}, {sigma, theta, nu}}; }
// CIR time change of a Levy base (test.cpp:887-905); GAUSS base = the reference's Heston
inline CF timeChanged(CF base, double speed, double adaV, double rho, double v0Hat,
                      bool diffusionOnly = false) {
  base.model.time_change = diffusionOnly ? FOP_TC_CIR_DIFFUSION : FOP_TC_CIR;
  base.params.insert(base.params.end(), {speed, adaV, rho, v0Hat});
  return base;
}
inline CF heston(double sigma, double speed, double adaV, double rho, double v0Hat) {
  return timeChanged(blackScholes(sigma), speed, adaV, rho, v0Hat);
}
enum class Payoff { Call = FOP_PAYOFF_CALL, Put = FOP_PAYOFF_PUT };

}  // namespace fftop

namespace optionprice {

// ---- helpers with the reference's names (OptionPricing.h:19-42, 84-138)
template <typename Index, typename Number>
auto getUDomain(const Number& uMax, const Number& du, const Index& index) {
  auto myU = du * index;
  return myU > uMax ? myU - 2.0 * uMax : myU;
}
template <typename Number> auto getMaxK(const Number& ada) { return M_PI / ada; }
template <typename Index, typename Number>
auto getLambda(const Index& numSteps, const Number& b) { return 2.0 * b / numSteps; }

inline std::vector<double> getStrikeUnderlying(double xMin, double xMax, double K, int numX) {
  std::vector<double> v(numX);
  fop_strike_underlying(xMin, xMax, K, numX, v.data());
  return v;
}
inline std::vector<double> getFangOostStrike(double xMin, double xMax, double S0, int numX) {
  std::vector<double> v(numX);
  fop_fang_oost_strikes(xMin, xMax, S0, numX, v.data());
  return v;
}
inline std::vector<double> getCarrMadanStrikes(double ada, double S0, int numX) {
  std::vector<double> v(numX);
  fop_carr_madan_strikes(numX, ada, S0, v.data());
  return v;
}

// getFangOostKAtIndex (OptionPricing.h:93-97), getXFromK (:110-117), getCarrMadanKAtIndex (:121-123)
inline double getFangOostKAtIndex(double xMin, double dK, double S0, int index) {
  return S0 * std::exp(-(xMin + dK * index));
}
template <typename Array>
Array getXFromK(double S0, const Array& K) {
  std::vector<double> x;
  for (const auto& k : K) x.push_back(std::log(S0 / k));
  return Array(x.begin(), x.end());
}
inline double getCarrMadanKAtIndex(double b, double lambda, double S0, int index) {
  return S0 * std::exp(-b + lambda * index);
}

// ---- Carr-Madan (OptionPricing.h:620-704)
inline std::vector<double> CarrMadanCall(int numSteps, double ada, double alpha, double S0, double r,
                                         double T, const fftop::CF& cf) {
  auto& e = fftop::Engine::instance();
  std::vector<double> out(numSteps);
  e.check(fop_carr_madan(e.get(), cf.model, cf.params.data(), 1, numSteps, ada, alpha, S0, r, T, 0,
                         out.data()));
  return out;
}
inline std::vector<double> CarrMadanPut(int numSteps, double ada, double alpha, double S0, double r,
                                        double T, const fftop::CF& cf) {
  auto& e = fftop::Engine::instance();
  std::vector<double> out(numSteps);
  e.check(fop_carr_madan(e.get(), cf.model, cf.params.data(), 1, numSteps, ada, alpha, S0, r, T, 1,
                         out.data()));
  return out;
}
inline std::vector<double> CarrMadanCall(int numSteps, double ada, double S0, double r, double T,
                                         const fftop::CF& cf) {
  return CarrMadanCall(numSteps, ada, 1.5, S0, r, T, cf);  // alpha = 1.5, OptionPricing.h:679
}
inline std::vector<double> CarrMadanPut(int numSteps, double ada, double S0, double r, double T,
                                        const fftop::CF& cf) {
  return CarrMadanPut(numSteps, ada, 1.5, S0, r, T, cf);
}

// ---- Fang-Oosterlee COS (OptionPricing.h:358-559): returns the caller's strike container type
namespace detail {
template <typename Array>
Array cosCall(int kind, double S0, const Array& K, double r, double T, int numU,
              const fftop::CF& cf) {
  auto& e = fftop::Engine::instance();
  std::vector<double> k(K.begin(), K.end()), out(k.size());
  e.check(fop_cos(e.get(), cf.model, cf.params.data(), 1, k.data(), (int)k.size(), S0, r, &T, 1,
                  numU, kind, out.data()));
  return Array(out.begin(), out.end());
}
}  // namespace detail
#define FFTOP_COS_ENTRY(NAME, KIND)                                                          \
  template <typename Array>                                                                  \
  Array NAME(double S0, const Array& K, double r, double T, int numUSteps, const fftop::CF& cf) { \
    return detail::cosCall(KIND, S0, K, r, T, numUSteps, cf);                                \
  }
FFTOP_COS_ENTRY(FangOostCallPrice, FOP_CALL_PRICE)
FFTOP_COS_ENTRY(FangOostPutPrice, FOP_PUT_PRICE)
FFTOP_COS_ENTRY(FangOostCallDelta, FOP_CALL_DELTA)
FFTOP_COS_ENTRY(FangOostPutDelta, FOP_PUT_DELTA)
FFTOP_COS_ENTRY(FangOostCallGamma, FOP_CALL_GAMMA)
FFTOP_COS_ENTRY(FangOostPutGamma, FOP_PUT_GAMMA)
FFTOP_COS_ENTRY(FangOostCallTheta, FOP_CALL_THETA)
FFTOP_COS_ENTRY(FangOostPutTheta, FOP_PUT_THETA)
#undef FFTOP_COS_ENTRY

// ---- FSTS (OptionPricing.h:185-281); steps/american extend it per README.md:19
namespace detail {
inline std::vector<double> fsts(int kind, int numSteps, double xmax, double r, double T,
                                double strike, fftop::Payoff payoff, const fftop::CF& cf,
                                int steps = 1, bool american = false) {
  auto& e = fftop::Engine::instance();
  std::vector<double> out(numSteps);
  e.check(fop_fsts(e.get(), cf.model, cf.params.data(), 1, numSteps, xmax, strike, r, T, steps,
                   american ? 1 : 0, (int)payoff, kind, out.data()));
  return out;
}
}  // namespace detail
inline std::vector<double> FSTSPrice(int numSteps, double xmax, double r, double T, double strike,
                                     fftop::Payoff payoff, const fftop::CF& cf, int steps = 1,
                                     bool american = false) {
  return detail::fsts(FOP_FSTS_PRICE, numSteps, xmax, r, T, strike, payoff, cf, steps, american);
}
inline std::vector<double> FSTSDelta(int numSteps, double xmax, double r, double T, double strike,
                                     fftop::Payoff payoff, const fftop::CF& cf) {
  return detail::fsts(FOP_FSTS_DELTA, numSteps, xmax, r, T, strike, payoff, cf);
}
inline std::vector<double> FSTSGamma(int numSteps, double xmax, double r, double T, double strike,
                                     fftop::Payoff payoff, const fftop::CF& cf) {
  return detail::fsts(FOP_FSTS_GAMMA, numSteps, xmax, r, T, strike, payoff, cf);
}
inline std::vector<double> FSTSTheta(int numSteps, double xmax, double r, double T, double strike,
                                     fftop::Payoff payoff, const fftop::CF& cf) {
  return detail::fsts(FOP_FSTS_THETA, numSteps, xmax, r, T, strike, payoff, cf);
}

// ---- the reference's own signatures with an arbitrary callable characteristic function ---------
// `cf` is any callable std::complex<double> -> std::complex<double> (what the reference's templates
// take); it is sampled here, on the host, at the points the pricer needs.
namespace detail {
template <typename F> struct is_descriptor : std::is_same<typename std::decay<F>::type, fftop::CF> {};
template <typename F> using callable_t = typename std::enable_if<!is_descriptor<F>::value, int>::type;

template <typename CFn>
std::vector<double> carrMadanSampled(int numSteps, double ada, double alpha, double S0, double r, double T,
                                     CFn&& cf, int isPut) {
  auto& e = fftop::Engine::instance();
  std::vector<double> u(2 * (size_t)numSteps), phi(2 * (size_t)numSteps), out(numSteps);
  fop_carr_madan_u_grid(numSteps, ada, alpha, u.data());
  for (int j = 0; j < numSteps; ++j) {
    const std::complex<double> v = cf(std::complex<double>(u[2 * j], u[2 * j + 1]));
    phi[2 * j] = v.real();
    phi[2 * j + 1] = v.imag();
  }
  e.check(fop_carr_madan_cf_samples(e.get(), phi.data(), 1, numSteps, ada, alpha, S0, r, T, isPut, out.data()));
  return out;
}
template <typename Array, typename CFn>
Array cosSampled(int kind, double S0, const Array& K, double r, double T, int numU, CFn&& cf) {
  auto& e = fftop::Engine::instance();
  std::vector<double> k(K.begin(), K.end()), out(k.size()), u(numU), phi(2 * (size_t)numU);
  fop_cos_u_grid(k.data(), (int)k.size(), S0, numU, u.data());
  for (int i = 0; i < numU; ++i) {
    const std::complex<double> v = cf(std::complex<double>(0.0, u[i]));
    phi[2 * i] = v.real();
    phi[2 * i + 1] = v.imag();
  }
  e.check(fop_cos_cf_samples(e.get(), phi.data(), 1, k.data(), (int)k.size(), S0, r, &T, 1, numU, kind, out.data()));
  return Array(out.begin(), out.end());
}
template <typename GetAssetPrice, typename PayoffFn, typename CFn>
std::vector<double> fstsSampled(int kind, int numSteps, double xmax, double r, double T, GetAssetPrice&& getAssetPrice,
                                PayoffFn&& payoff, CFn&& cf, int steps = 1, bool american = false) {
  auto& e = fftop::Engine::instance();
  const double dx = 2.0 * xmax / (numSteps - 1.0);  // fangoost::computeDX(numSteps, -xmax, xmax)
  std::vector<double> w(numSteps), phi(2 * (size_t)numSteps), pay(numSteps), asset(numSteps), out(numSteps);
  fop_fsts_u_grid(numSteps, xmax, w.data());
  for (int j = 0; j < numSteps; ++j) {
    asset[j] = getAssetPrice(-xmax + dx * j);
    pay[j] = payoff(asset[j]);
    const std::complex<double> v = cf(std::complex<double>(0.0, w[j]));
    phi[2 * j] = v.real();
    phi[2 * j + 1] = v.imag();
  }
  e.check(fop_fsts_samples(e.get(), phi.data(), pay.data(), 0, asset.data(), 1, numSteps, xmax, r, T, steps,
                           american ? 1 : 0, kind, out.data()));
  return out;
}
}  // namespace detail

template <typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> CarrMadanCall(int numSteps, double ada, double alpha, double S0, double r, double T, CFn&& cf) {
  return detail::carrMadanSampled(numSteps, ada, alpha, S0, r, T, cf, 0);
}
template <typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> CarrMadanPut(int numSteps, double ada, double alpha, double S0, double r, double T, CFn&& cf) {
  return detail::carrMadanSampled(numSteps, ada, alpha, S0, r, T, cf, 1);
}
template <typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> CarrMadanCall(int numSteps, double ada, double S0, double r, double T, CFn&& cf) {
  return detail::carrMadanSampled(numSteps, ada, 1.5, S0, r, T, cf, 0);
}
template <typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> CarrMadanPut(int numSteps, double ada, double S0, double r, double T, CFn&& cf) {
  return detail::carrMadanSampled(numSteps, ada, 1.5, S0, r, T, cf, 1);
}
#define FFTOP_COS_CALLABLE(NAME, KIND)                                                       \
  template <typename Array, typename CFn, detail::callable_t<CFn> = 0>                       \
  Array NAME(double S0, const Array& K, double r, double T, int numUSteps, CFn&& cf) {        \
    return detail::cosSampled(KIND, S0, K, r, T, numUSteps, cf);                             \
  }
FFTOP_COS_CALLABLE(FangOostCallPrice, FOP_CALL_PRICE)
FFTOP_COS_CALLABLE(FangOostPutPrice, FOP_PUT_PRICE)
FFTOP_COS_CALLABLE(FangOostCallDelta, FOP_CALL_DELTA)
FFTOP_COS_CALLABLE(FangOostPutDelta, FOP_PUT_DELTA)
FFTOP_COS_CALLABLE(FangOostCallGamma, FOP_CALL_GAMMA)
FFTOP_COS_CALLABLE(FangOostPutGamma, FOP_PUT_GAMMA)
FFTOP_COS_CALLABLE(FangOostCallTheta, FOP_CALL_THETA)
FFTOP_COS_CALLABLE(FangOostPutTheta, FOP_PUT_THETA)
#undef FFTOP_COS_CALLABLE
// FSTSPrice(numSteps, xmax, r, T, getAssetPrice, payoff, cf) etc. -- the reference's argument list
// (OptionPricing.h:185-281); steps / american extend it (README.md:19)
template <typename G, typename Pf, typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> FSTSPrice(int numSteps, double xmax, double r, double T, G&& getAssetPrice, Pf&& payoff, CFn&& cf,
                              int steps = 1, bool american = false) {
  return detail::fstsSampled(FOP_FSTS_PRICE, numSteps, xmax, r, T, getAssetPrice, payoff, cf, steps, american);
}
template <typename G, typename Pf, typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> FSTSDelta(int numSteps, double xmax, double r, double T, G&& getAssetPrice, Pf&& payoff, CFn&& cf) {
  return detail::fstsSampled(FOP_FSTS_DELTA, numSteps, xmax, r, T, getAssetPrice, payoff, cf);
}
template <typename G, typename Pf, typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> FSTSGamma(int numSteps, double xmax, double r, double T, G&& getAssetPrice, Pf&& payoff, CFn&& cf) {
  return detail::fstsSampled(FOP_FSTS_GAMMA, numSteps, xmax, r, T, getAssetPrice, payoff, cf);
}
template <typename G, typename Pf, typename CFn, detail::callable_t<CFn> = 0>
std::vector<double> FSTSTheta(int numSteps, double xmax, double r, double T, G&& getAssetPrice, Pf&& payoff, CFn&& cf) {
  return detail::fstsSampled(FOP_FSTS_THETA, numSteps, xmax, r, T, getAssetPrice, payoff, cf);
}

}  // namespace optionprice
#endif
