// Drop-in C++ host header for the reference's fft.h / fft.hpp over the B200 engine: fft, ifft
// (fft.hpp:49-61, unnormalised, moved through &&) and integrateFFT / integrateIFFT (fft.hpp:64-110).
// The integrand stays an arbitrary host callable, as in the reference: it is sampled on the host at
// the N grid points and the weighted transform runs on the device (fop_integrate_fft).
#ifndef FFTOP_FFT_H
#define FFTOP_FFT_H
#include <complex>
#include <utility>
#include <vector>

#include "OptionPricing.h"

typedef std::complex<double> Complex;
typedef std::vector<Complex> CArray;

inline CArray fft(CArray&& x) {
  auto& e = fftop::Engine::instance();
  e.check(fop_fft(e.get(), reinterpret_cast<double*>(x.data()), 1, (int)x.size(), 0));
  return std::move(x);
}
inline CArray ifft(CArray&& x) {
  auto& e = fftop::Engine::instance();
  e.check(fop_fft(e.get(), reinterpret_cast<double*>(x.data()), 1, (int)x.size(), 1));
  return std::move(x);
}

namespace fftop {
template <typename FN, typename EditOutput>
CArray integrate(double outputMin, double inputMin, double inputMax, FN&& fn, int N, int inverse,
                 EditOutput&& fnOutput) {
  CArray s(N);
  const double inputDx = (inputMax - inputMin) / (N - 1);
  for (int i = 0; i < N; ++i) s[i] = fn(inputMin + inputDx * i);
  auto& e = Engine::instance();
  e.check(fop_integrate_fft(e.get(), reinterpret_cast<const double*>(s.data()), 1, N, outputMin,
                            inputMin, inputMax, inverse, reinterpret_cast<double*>(s.data())));
  const double du = ((N - 1) * 2.0 * M_PI / (inputMax - inputMin)) / N;
  for (int i = 0; i < N; ++i) s[i] = fnOutput(outputMin + i * du, s[i]);
  return s;
}
struct Identity {
  Complex operator()(double, const Complex& v) const { return v; }
};
}  // namespace fftop

template <typename FN>
CArray integrateFFT(double xMin, double uMin, double uMax, FN&& fn, int N) {
  return fftop::integrate(xMin, uMin, uMax, fn, N, 0, fftop::Identity());
}
template <typename FN, typename EditOutput>
CArray integrateFFT(double xMin, double uMin, double uMax, FN&& fn, int N, EditOutput&& fnOutput) {
  return fftop::integrate(xMin, uMin, uMax, fn, N, 0, fnOutput);
}
template <typename FN>
CArray integrateIFFT(double uMin, double xMin, double xMax, FN&& fn, int N) {
  return fftop::integrate(uMin, xMin, xMax, fn, N, 1, fftop::Identity());
}
template <typename FN, typename EditOutput>
CArray integrateIFFT(double uMin, double xMin, double xMax, FN&& fn, int N, EditOutput&& fnOutput) {
  return fftop::integrate(uMin, xMin, xMax, fn, N, 1, fnOutput);
}
#endif
