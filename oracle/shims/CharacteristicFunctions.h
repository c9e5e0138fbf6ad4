// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Shim for the un-vendored third-party header `CharacteristicFunctions.h`
// (phillyfan1138/CharacteristicFunctions @0dfefb3dc845d40ae94b70edb23cf6301bd6df18,
// pinned at /root/reference/package.sh:21).  Source NOT in /root/reference.  Formulas
// restated from the model write-up docs/calibration/OptionCalibration.Rnw:25-37 and from
// the argument orders at the reference's call sites:
//   gaussCF(u, mu, sigma)                        test.cpp:54
//   gaussLogCF(u, mu, sigma)                     test.cpp:570, :898
//   mertonLogRNCF(u, lambda, muL, sigL, r, sig)  test.cpp:1087-1090, :1119
//   cgmyLogRNCF(u, C, G, M, Y, r, sigma)         test.cpp:731, :888
//   cirLogMGF(psi, a, kappa, sigma, t, v0)       test.cpp:887-894 (kappa may be complex)
// Every `u` is the complex number i*omega (or 1+i*omega in the calibration objective).
// Complex sqrt/log/pow are the principal branches of libstdc++'s std::complex.
// Pinned by the reference's constants 108.445317144 (test.cpp:825), 108.4614527579 (:862),
// 72.9703833045 (:953), 74.9251/75.0101 (:951-952), 5.78515545 (:1027,:1065), 5.9713 (:1127).
#ifndef FFTOP_ORACLE_SHIM_CHARACTERISTICFUNCTIONS_H
#define FFTOP_ORACLE_SHIM_CHARACTERISTICFUNCTIONS_H
#include <complex>
#include <cmath>

namespace chfunctions {

template <typename U, typename Mu, typename Sigma>
auto gaussLogCF(const U& u, const Mu& mu, const Sigma& sigma) {
  return u * mu + 0.5 * sigma * sigma * u * u;
}
template <typename U, typename Mu, typename Sigma>
auto gaussCF(const U& u, const Mu& mu, const Sigma& sigma) {
  return exp(gaussLogCF(u, mu, sigma));
}

template <typename U, typename Number>
auto mertonLogCF(const U& u, const Number& lambda, const Number& muL, const Number& sigL,
                 const Number& mu, const Number& sigma) {
  return gaussLogCF(u, mu, sigma) + lambda * (gaussCF(u, muL, sigL) - 1.0);
}
// risk-neutral drift: mu = r - sigma^2/2 - lambda (E[e^J] - 1)
template <typename U, typename Number>
auto mertonLogRNCF(const U& u, const Number& lambda, const Number& muL, const Number& sigL,
                   const Number& r, const Number& sigma) {
  return mertonLogCF(u, lambda, muL, sigL,
                     r - sigma * sigma * 0.5 -
                         lambda * (std::exp(muL + 0.5 * sigL * sigL) - 1.0),
                     sigma);
}

// C Gamma(-Y) [ (M-u)^Y + (G+u)^Y - M^Y - G^Y ];  identically 0 when C == 0 (test.cpp:783)
template <typename U, typename Number>
auto cgmyLogCF(const U& u, const Number& C, const Number& G, const Number& M,
               const Number& Y) {
  return C == 0.0 ? U(0.0)
                  : U(C * std::tgamma(-Y) *
                      (pow(M - u, Y) + pow(G + u, Y) - std::pow(M, Y) - std::pow(G, Y)));
}
template <typename U, typename Number>
auto cgmyLogRNCF(const U& u, const Number& C, const Number& G, const Number& M,
                 const Number& Y, const Number& r, const Number& sigma) {
  return gaussLogCF(u, r - sigma * sigma * 0.5 - cgmyLogCF(1.0, C, G, M, Y), sigma) +
         cgmyLogCF(u, C, G, M, Y);
}

// Variance gamma (BASELINE.json north_star names "CGMY/VG").  NOT in the reference and not in the
// pinned CharacteristicFunctions revision's call sites: parity UNPINNED -- textbook form (Madan,
// Carr, Chang 1998), X_t = theta G_t + sigma W(G_t), G a gamma process of variance rate nu:
//   E e^{u X_t} = (1 - u theta nu - sigma^2 nu u^2 / 2)^(-t/nu)
// checked against a 40-digit evaluation in tests/mp_exact.py.
template <typename U, typename Number>
auto vgLogCF(const U& u, const Number& sigma, const Number& theta, const Number& nu) {
  return -(1.0 / nu) * log(1.0 - u * theta * nu - 0.5 * sigma * sigma * nu * u * u);
}
template <typename U, typename Number>
auto vgLogRNCF(const U& u, const Number& sigma, const Number& theta, const Number& nu, const Number& r) {
  return u * (r + std::log(1.0 - theta * nu - 0.5 * sigma * sigma * nu) / nu) + vgLogCF(u, sigma, theta, nu);
}

// log E[exp(-psi int_0^t v_s ds)],  dv = (a - kappa v) dt + sigma sqrt(v) dW.
template <typename Psi, typename A, typename Kappa, typename Sigma, typename T, typename V0>
auto cirLogMGF(const Psi& psi, const A& a, const Kappa& kappa, const Sigma& sigma,
               const T& t, const V0& v0) {
  using C = std::complex<double>;
  const C ps(psi), kp(kappa);
  if (sigma == 0.0) {  // deterministic time change (test.cpp:1187)
    const C ek = 1.0 - exp(-kp * t);
    const C b = ps * ek / kp;
    const C c = (a * ps / kp) * (t - ek / kp);
    return -b * v0 - c;
  }
  const C delta = sqrt(kp * kp + 2.0 * ps * sigma * sigma);
  const C e = exp(-delta * t);
  const C dmk = delta - kp;
  const C b = 2.0 * ps * (1.0 - e) / (delta + kp + dmk * e);
  const C c = (a / (sigma * sigma)) *
              (2.0 * log(1.0 - dmk * (1.0 - e) / (2.0 * delta)) + dmk * t);
  return -b * v0 - c;
}

}  // namespace chfunctions
#endif
