// ODE-based characteristic function of the reference's calibration driver
// (generateCharts.cpp:391-439, `cfLogGeneric`): a time-changed Merton jump diffusion whose variance
// process jumps together with the asset (self-exciting volatility jumps of exponential size), an
// affine model whose log CF solves the Duffie-Pan-Singleton Riccati system
//     x1' = beta(x1, cfPart),   x2' = alpha(x1, cfPart),   x(0) = (0, 0),
//     log phi(u) = x1(T) v0 + x2(T),
// integrated with classical Runge-Kutta (numODE = 64 steps, generateCharts.cpp:403).
//
// The reference builds it from un-vendored pieces -- chfunctions::AlphaOrBeta_move,
// exponentialCFBeta, logAffine (CharacteristicFunctions) and rungekutta::compute_efficient_2d
// (RungeKutta @6326974b, package.sh:25) -- and never asserts a value for it, so this restatement is
// parity-UNPINNED.  What pins the conventions: with lambda = 0 the system must reduce to the CIR
// time change the reference's own tests pin (cirLogMGF, test.cpp:1033-1068), i.e.
//     x1' = psi(u) - (speed - u rho sigma adaV) x1 + adaV^2 x1^2 / 2,   x2' = speed x1,
// which fixes AlphaOrBeta(rho, K, H, l)(x, c) = -rho + K x + H x^2 / 2 + l c; tests/cpp/test_dropin.cpp
// checks exactly that against the on-device Heston model.
//
// Host code: the result is an ordinary callable, priced on the GPU through the callable overloads
// of OptionPricing.h (fop_cos_cf_samples / fop_carr_madan_cf_samples / fop_fsts_samples).
#ifndef FFTOP_AFFINEODE_H
#define FFTOP_AFFINEODE_H
#include <array>
#include <complex>

namespace chfunctions {
typedef std::complex<double> Cplx;

// -rho + K x + H x^2 / 2 + l cfPart   (the Riccati right-hand sides of an affine jump diffusion)
struct AlphaOrBeta {
  Cplx rho, K, H, l;
  Cplx operator()(const Cplx& x, const Cplx& cfPart) const { return -rho + K * x + 0.5 * H * x * x + l * cfPart; }
};
template <typename A, typename B, typename C, typename D>
AlphaOrBeta AlphaOrBeta_move(const A& rho, const B& K, const C& H, const D& l) {
  return AlphaOrBeta{Cplx(rho), Cplx(K), Cplx(H), Cplx(l)};
}
// moment generating function of an exponential jump of mean q
inline Cplx exponentialCFBeta(const Cplx& u, double q) { return 1.0 / (1.0 - q * u); }
inline Cplx gaussCFHost(const Cplx& u, double mu, double sigma) { return std::exp(u * mu + 0.5 * sigma * sigma * u * u); }
inline Cplx mertonLogRNCFHost(const Cplx& u, double lambda, double muL, double sigL, double r, double sigma) {
  const double mu = r - 0.5 * sigma * sigma - lambda * (std::exp(muL + 0.5 * sigL * sigL) - 1.0);
  return u * mu + 0.5 * sigma * sigma * u * u + lambda * (gaussCFHost(u, muL, sigL) - 1.0);
}
inline Cplx logAffine(const std::array<Cplx, 2>& x, double v0) { return x[0] * v0 + x[1]; }
}  // namespace chfunctions

namespace rungekutta {
// classical fourth-order Runge-Kutta for a system of two complex ODEs, numSteps equal steps on [0, T]
template <typename F>
std::array<std::complex<double>, 2> compute_efficient_2d(double T, int numSteps, std::array<std::complex<double>, 2> x, F&& fn) {
  typedef std::complex<double> C;
  const double h = T / numSteps;
  for (int i = 0; i < numSteps; ++i) {
    const double t = h * i;
    const std::array<C, 2> k1 = fn(t, x[0], x[1]);
    const std::array<C, 2> k2 = fn(t + 0.5 * h, x[0] + 0.5 * h * k1[0], x[1] + 0.5 * h * k1[1]);
    const std::array<C, 2> k3 = fn(t + 0.5 * h, x[0] + 0.5 * h * k2[0], x[1] + 0.5 * h * k2[1]);
    const std::array<C, 2> k4 = fn(t + h, x[0] + h * k3[0], x[1] + h * k3[1]);
    x[0] += h / 6.0 * (k1[0] + 2.0 * k2[0] + 2.0 * k3[0] + k4[0]);
    x[1] += h / 6.0 * (k1[1] + 2.0 * k2[1] + 2.0 * k3[1] + k4[1]);
  }
  return x;
}
}  // namespace rungekutta

namespace fftop {
// cfLogGeneric(T)(lambda, muJ, sigJ, sigma, v0, speed, adaV, rho, delta) -> u -> log phi(u)
// (generateCharts.cpp:391-439, argument order kept)
struct CfLogGeneric {
  double T, lambda, muJ, sigJ, sigma, v0, speed, adaV, rho, delta;
  int numODE = 64;
  std::complex<double> operator()(const std::complex<double>& u) const {
    typedef std::complex<double> C;
    const chfunctions::AlphaOrBeta alpha = chfunctions::AlphaOrBeta_move(0.0, speed, 0.0, 0.0);
    const chfunctions::AlphaOrBeta beta = chfunctions::AlphaOrBeta_move(
        -chfunctions::mertonLogRNCFHost(u, lambda, muJ, sigJ, 0.0, sigma),
        -(speed + delta * lambda - u * rho * sigma * adaV), adaV * adaV, lambda);
    const C jumpS = chfunctions::gaussCFHost(u, muJ, sigJ);
    const double dl = delta;
    return chfunctions::logAffine(
        rungekutta::compute_efficient_2d(T, numODE, std::array<C, 2>{C(0.0, 0.0), C(0.0, 0.0)},
                                         [&](double, const C& x1, const C&) {
                                           const C cfPart = (chfunctions::exponentialCFBeta(x1, dl) - 1.0) * jumpS;
                                           return std::array<C, 2>{beta(x1, cfPart), alpha(x1, cfPart)};
                                         }),
        v0);
  }
};
inline auto cfLogGeneric(double T) {
  return [T](double lambda, double muJ, double sigJ, double sigma, double v0, double speed, double adaV, double rho,
             double delta) { return CfLogGeneric{T, lambda, muJ, sigJ, sigma, v0, speed, adaV, rho, delta}; };
}
}  // namespace fftop
#endif
