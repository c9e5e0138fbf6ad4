// The reference's own test cases (test.cpp) re-expressed against the drop-in headers.
// Built and run by tests/test_cpp_dropin_gpu.py on a GPU box.
#include <cmath>
#include <complex>
#include <cstdio>
#include <deque>
#include <vector>

#include "fftop/AffineODE.h"
#include "fftop/OptionCalibration.h"
#include "fftop/OptionPricing.h"
#include "fftop/cuckoo.h"
#include "fftop/fft.h"

static int failures = 0;
// Catch v1 Approx semantics: |a-b| < eps (1 + max(|a|,|b|))
#define REQUIRE_APPROX(a, b, eps)                                                       \
  do {                                                                                  \
    const double a_ = (a), b_ = (b);                                                    \
    if (!(std::fabs(a_ - b_) < (eps) * (1.0 + std::fmax(std::fabs(a_), std::fabs(b_))))) { \
      std::printf("FAIL %s:%d  %.12g vs %.12g\n", __FILE__, __LINE__, a_, b_);          \
      ++failures;                                                                       \
    }                                                                                   \
  } while (0)

// test.cpp:13-21
static double BSCall(double S0, double discount, double k, double sigma, double T) {
  const double s = std::sqrt(2.0), ss = std::sqrt(T) * sigma;
  const double d1 = std::log(S0 / (discount * k)) / ss + ss * .5;
  return S0 * (.5 + .5 * std::erf(d1 / s)) - k * discount * (.5 + .5 * std::erf((d1 - ss) / s));
}

int main() {
  const double r = .05, sig = .3, T = 1.0, S0 = 50.0, discount = std::exp(-r * T);
  {  // CarrMadanCall, test.cpp:619-648
    const int numX = 1024;
    const double ada = .25;
    auto price = optionprice::CarrMadanCall(numX, ada, S0, r, T, fftop::blackScholes(sig));
    auto K = optionprice::getCarrMadanStrikes(ada, S0, numX);
    for (int i = (int)(numX * .3); i < (int)(numX * .7); ++i)
      REQUIRE_APPROX(price[i], BSCall(S0, discount, K[i], sig, T), 1e-4);
  }
  {  // FangOosterleeCallFewKLowT with a std::deque, test.cpp:562-588
    std::deque<double> K = {7500, 60, 50, 40, .3};
    const double Tl = .25;
    auto price = optionprice::FangOostCallPrice(S0, K, r, Tl, 256, fftop::blackScholes(sig));
    for (int i = 1; i < 4; ++i) REQUIRE_APPROX(price[i], BSCall(S0, std::exp(-r * Tl), K[i], sig, Tl), 1e-4);
  }
  {  // FSTSCall, test.cpp:47-79
    const int numX = 1024;
    const double xmax = 5.0, K = 50.0;
    auto price = optionprice::FSTSPrice(numX, xmax, r, T, K, fftop::Payoff::Call, fftop::blackScholes(sig));
    auto S = optionprice::getStrikeUnderlying(-xmax, xmax, K, numX);
    for (int i = (int)(numX * .3); i < (int)(numX * .7); ++i)
      REQUIRE_APPROX(price[i], BSCall(S[i], discount, K, sig, T), 1e-4);
  }
  {  // FangOosterleePutCGMYWithVol, test.cpp:799-829 (12-digit constant)
    std::vector<double> K = {7500, 500.0, .3};
    auto p = optionprice::FangOostPutPrice(500.0, K, .4, .25, 256, fftop::cgmy(.2, 1.0, 1.4, 2.5, 1.4));
    REQUIRE_APPROX(p[1], 108.445317144, 1e-10);
  }
  {  // CarrMadanCGMY, test.cpp:831-865
    auto p = optionprice::CarrMadanPut(1024, .25, 500.0, .4, .25, fftop::cgmy(.2, 1.0, 1.4, 2.5, 1.4));
    REQUIRE_APPROX(p[512], 108.4614527579, 1e-10);
  }
  {  // FangOostHeston, test.cpp:1033-1068
    const double b = .0398, a = 1.5768, c = .5751, rho = -.5711, v0 = .0175;
    std::vector<double> K = {5000, 100, .3};
    auto p = optionprice::FangOostCallPrice(100.0, K, 0.0, 1.0, 256,
                                            fftop::heston(std::sqrt(b), a, c / std::sqrt(b), rho, v0 / b));
    REQUIRE_APPROX(p[1], 5.78515545, 1e-4);
  }
  {  // getObjFn_arr: NaN -> 5000 penalty (OptionCalibration.h:185,196) and exactness at the target
    std::vector<double> u = {6, 7, 8};
    std::vector<std::complex<double>> ph(3, std::complex<double>(0.0, 0.0));
    auto hoc = optioncal::getObjFn_arr(std::move(ph), fftop::blackScholes(.3), std::move(u), 0.0, 1.0);
    std::vector<double> bad = {std::nan("")};
    REQUIRE_APPROX(hoc(bad), 5000.0, 1e-15);
  }
  {  // cuckoo::optimize call shape of generateCharts.cpp:102-118 (BS: one parameter, target sigma = .3)
    std::vector<double> u = {-15, -10, -5, 5, 10, 15};
    std::vector<std::complex<double>> ph;
    for (double ul : u) {  // log CF of BS at 1 + i u: (1+iu)(r - s^2/2) T + s^2 (1+iu)^2 T / 2, r = 0, T = 1
      const std::complex<double> z(1.0, ul);
      ph.push_back(z * (-.045) + .045 * z * z);
    }
    auto objFn = optioncal::getObjFn_arr(std::move(ph), fftop::blackScholes(.3), std::move(u), 0.0, 1.0);
    std::vector<swarm_utils::upper_lower<double>> constraints = {swarm_utils::upper_lower<double>(0.0, .6)};
    const auto results = cuckoo::optimize(objFn, constraints, 25, 300, 1e-20, 42);
    auto params = std::get<swarm_utils::optparms>(results);
    auto objFnRes = std::get<swarm_utils::fnval>(results);
    REQUIRE_APPROX(params[0], .3, 1e-4);
    REQUIRE_APPROX(objFnRes, objFn(params), 1e-15);
  }
  {  // fft.h drop-in: ifft(fft(x)) = N x (fft.hpp:1-4: no 1/N either way) and integrateFFT of a
     // normal density = its characteristic function (frequency -m du inside the sum, fft.hpp:70-81)
    const int N = 1024;
    CArray x(N);
    for (int i = 0; i < N; ++i) x[i] = Complex(std::sin(0.1 * i), std::cos(0.3 * i));
    CArray y = ifft(fft(CArray(x)));
    for (int i = 0; i < N; i += 37) {
      REQUIRE_APPROX(y[i].real() / N, x[i].real(), 1e-12);
      REQUIRE_APPROX(y[i].imag() / N, x[i].imag(), 1e-12);
    }
    const double lo = -8.0, hi = 8.0, mu = .3, sg = .8;
    auto dens = [&](double t) { return Complex(std::exp(-.5 * (t - mu) * (t - mu) / (sg * sg)) / (sg * std::sqrt(2 * M_PI)), 0.0); };
    CArray f = integrateFFT(0.0, lo, hi, dens, 4096);
    const double du = ((4096 - 1) * 2.0 * M_PI / (hi - lo)) / 4096;
    for (int m = 0; m < 30; m += 7) {
      const double v = -m * du;
      const Complex expect = std::exp(Complex(-.5 * sg * sg * v * v, v * mu)) * std::exp(Complex(0.0, -v * lo)) *
                             std::exp(Complex(0.0, m * du * lo));
      REQUIRE_APPROX(f[m].real(), expect.real(), 1e-9);
      REQUIRE_APPROX(f[m].imag(), expect.imag(), 1e-9);
    }
  }
  {  // the reference's OWN call shapes with arbitrary callables (test.cpp:47-79, 342-367, 619-648):
     // lambdas for the characteristic function, the payoff and the asset map
    typedef std::complex<double> C;
    const int numX = 1024;
    const double xmax = 5.0, K = 50.0, ada = .25;
    auto BSCF = [&](const C& u) { return std::exp(r * T * u + sig * sig * .5 * T * (u * u - u)); };  // gaussCF, test.cpp:54
    auto payoff = [&](double assetValue) { return assetValue > K ? assetValue - K : 0.0; };           // test.cpp:56-58
    auto getAssetPrice = [&](double logAsset) { return K * std::exp(logAsset); };                    // test.cpp:59-61
    auto fsts = optionprice::FSTSPrice(numX, xmax, r, T, getAssetPrice, payoff, BSCF);
    auto fstsD = optionprice::FSTSPrice(numX, xmax, r, T, K, fftop::Payoff::Call, fftop::blackScholes(sig));
    auto S = optionprice::getStrikeUnderlying(-xmax, xmax, K, numX);
    for (int i = (int)(numX * .3); i < (int)(numX * .7); ++i) {
      REQUIRE_APPROX(fsts[i], BSCall(S[i], discount, K, sig, T), 1e-4);
      REQUIRE_APPROX(fsts[i], fstsD[i], 1e-9);     // sampled lambda == on-device model
    }
    auto dl = optionprice::FSTSDelta(numX, xmax, r, T, getAssetPrice, payoff, BSCF);
    auto dlD = optionprice::FSTSDelta(numX, xmax, r, T, K, fftop::Payoff::Call, fftop::blackScholes(sig));
    for (int i = (int)(numX * .3); i < (int)(numX * .7); ++i) REQUIRE_APPROX(dl[i], dlD[i], 1e-9);
    // American put through lambdas: 64 steps of the one-step CF
    const int steps = 64;
    auto BSCFdt = [&](const C& u) { return std::exp((r * u + sig * sig * .5 * (u * u - u)) * (T / steps)); };
    auto putPay = [&](double a) { return a < K ? K - a : 0.0; };
    auto am = optionprice::FSTSPrice(numX, xmax, r, T, getAssetPrice, putPay, BSCFdt, steps, true);
    auto amD = optionprice::FSTSPrice(numX, xmax, r, T, K, fftop::Payoff::Put, fftop::blackScholes(sig), steps, true);
    for (int i = (int)(numX * .3); i < (int)(numX * .7); ++i) REQUIRE_APPROX(am[i], amD[i], 1e-9);

    auto cm = optionprice::CarrMadanCall(numX, ada, S0, r, T, BSCF);
    auto cmD = optionprice::CarrMadanCall(numX, ada, S0, r, T, fftop::blackScholes(sig));
    for (int i = (int)(numX * .3); i < (int)(numX * .7); ++i) REQUIRE_APPROX(cm[i], cmD[i], 1e-9);

    std::deque<double> Kd = {7500, 60, 50, 40, .3};
    auto fo = optionprice::FangOostCallPrice(S0, Kd, r, T, 64, BSCF);
    auto foD = optionprice::FangOostCallPrice(S0, Kd, r, T, 64, fftop::blackScholes(sig));
    auto foG = optionprice::FangOostCallGamma(S0, Kd, r, T, 64, BSCF);
    auto foGD = optionprice::FangOostCallGamma(S0, Kd, r, T, 64, fftop::blackScholes(sig));
    for (int i = 1; i < 4; ++i) {
      REQUIRE_APPROX(fo[i], BSCall(S0, discount, Kd[i], sig, T), 1e-4);
      REQUIRE_APPROX(fo[i], foD[i], 1e-10);
      REQUIRE_APPROX(foG[i], foGD[i], 1e-10);
    }
    // helper names of the reference (OptionPricing.h:93-97, 110-117, 121-123)
    auto x = optionprice::getXFromK(S0, Kd);
    REQUIRE_APPROX(x[1], std::log(S0 / 60.0), 1e-15);
    REQUIRE_APPROX(optionprice::getFangOostKAtIndex(-5.0, 10.0 / 1023, S0, 512),
                   optionprice::getFangOostStrike(-5.0, 5.0, S0, 1024)[512], 1e-15);
    const double b = optionprice::getMaxK(ada), lambda = optionprice::getLambda(numX, b);
    REQUIRE_APPROX(optionprice::getCarrMadanKAtIndex(b, lambda, S0, 400), optionprice::getCarrMadanStrikes(ada, S0, numX)[400], 1e-15);
  }
  {  // ODE-based characteristic function of the calibration driver (generateCharts.cpp:391-439): Riccati
     // system solved by 64 Runge-Kutta steps on the host, priced on the GPU through the callable overload.
     // lambda = 0 must reduce to the CIR time change = the on-device Heston model (test.cpp:1033-1068).
    typedef std::complex<double> C;
    const double b = .0398, a = 1.5768, c = .5751, rho = -.5711, v0 = .0175;   // test.cpp:1035-1048
    const double sigma = std::sqrt(b), speed = a, adaV = c / std::sqrt(b), v0Hat = v0 / b, Tm = 1.0;
    const auto lcf = fftop::cfLogGeneric(Tm)(0.0, -.05, .1, sigma, v0Hat, speed, adaV, rho, .2);
    auto cf = [&](const C& u) { return std::exp(lcf(u)); };       // r = 0 as in the reference's Heston test
    std::vector<double> K = {5000, 120, 100, 80, .3};
    auto ode = optionprice::FangOostCallPrice(100.0, K, 0.0, Tm, 256, cf);
    auto dev = optionprice::FangOostCallPrice(100.0, K, 0.0, Tm, 256, fftop::heston(sigma, speed, adaV, rho, v0Hat));
    for (int i = 1; i < 4; ++i) REQUIRE_APPROX(ode[i], dev[i], 2e-7);     // RK4, 64 steps
    REQUIRE_APPROX(dev[2], 5.78515545, 1e-4);                              // test.cpp:1065
    // with volatility jumps the model stays a martingale: phi(1) = 1 (r = 0), and prices move
    const auto lj = fftop::cfLogGeneric(Tm)(.3, -.05, .1, sigma, v0Hat, speed, adaV, rho, .2);
    REQUIRE_APPROX(std::abs(std::exp(lj(C(1.0, 0.0)))), 1.0, 1e-6);
    auto cfj = [&](const C& u) { return std::exp(lj(u)); };
    auto odej = optionprice::FangOostCallPrice(100.0, K, 0.0, Tm, 256, cfj);
    if (!(odej[2] > ode[2] + 1e-3)) { std::printf("FAIL jumps should add value: %g vs %g\n", odej[2], ode[2]); ++failures; }
  }
  std::printf(failures ? "FAILED %d\n" : "ALL PASSED\n", failures);
  return failures != 0;
}
