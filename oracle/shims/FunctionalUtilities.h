// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Shim for the un-vendored third-party header `FunctionalUtilities.h`
// (phillyfan1138/FunctionalUtilities @20bb2555ccb00539496860bf598eb466d348c043,
// pinned at /root/reference/package.sh:19).  Its source is NOT in /root/reference;
// the contract below is restated from the reference's own call sites:
//   for_each_parallel(begin,end,f)          OptionPricing.h:87,101,135,171,181,597,604  fft.hpp:70-71
//   for_each_parallel(container&&, f)       OptionPricing.h:169-179 (in place, same container type back)
//   for_each_parallel_copy(container, f)    OptionPricing.h:111-116
//   for_each_parallel_subset(c&&, s, e, f)  OptionCalibration.h:66-68 (pinned by test.cpp:1214-1222)
//   sum(container, f) / sum(begin,end,f)    OptionCalibration.h:48,194
// The real library parallelises with OpenMP (makefile:13 passes -fopenmp); so do we.
// Nested regions are serialised by the OpenMP runtime, which lets the CPU baseline
// put its own `omp parallel for` over parameter sets around single-threaded calls.
#ifndef FFTOP_ORACLE_SHIM_FUNCTIONALUTILITIES_H
#define FFTOP_ORACLE_SHIM_FUNCTIONALUTILITIES_H
#include <vector>
#include <utility>
#include <type_traits>

namespace futilities {

template <typename F>
auto for_each_parallel(int begin, int end, F&& f)
    -> std::vector<typename std::decay<decltype(f(begin))>::type> {
  using R = typename std::decay<decltype(f(begin))>::type;
  const int n = end > begin ? end - begin : 0;
  std::vector<R> out(n);
#pragma omp parallel for
  for (int i = 0; i < n; ++i) out[i] = f(begin + i);
  return out;
}

// In-place map: element i becomes f(element i, i); the (moved) container is returned.
template <typename Container, typename F,
          typename = typename std::enable_if<!std::is_arithmetic<
              typename std::decay<Container>::type>::value>::type>
auto for_each_parallel(Container&& c, F&& f) ->
    typename std::decay<Container>::type {
  const int n = static_cast<int>(c.size());
#pragma omp parallel for
  for (int i = 0; i < n; ++i) c[i] = f(c[i], i);
  return std::move(c);
}

template <typename Container, typename F>
auto for_each_parallel_copy(const Container& c, F&& f) -> Container {
  Container out(c);
  const int n = static_cast<int>(c.size());
#pragma omp parallel for
  for (int i = 0; i < n; ++i) out[i] = f(c[i], i);
  return out;
}

// Applies f to indices [startOffset, size-endOffset); the rest is left untouched.
template <typename Container, typename F>
auto for_each_parallel_subset(Container&& c, int startOffset, int endOffset, F&& f) ->
    typename std::decay<Container>::type {
  const int n = static_cast<int>(c.size()) - endOffset;
#pragma omp parallel for
  for (int i = startOffset; i < n; ++i) c[i] = f(c[i], i);
  return std::move(c);
}

template <typename Container, typename F,
          typename = typename std::enable_if<!std::is_arithmetic<
              typename std::decay<Container>::type>::value>::type>
auto sum(const Container& c, F&& f) ->
    typename std::decay<decltype(f(c[0], 0))>::type {
  using R = typename std::decay<decltype(f(c[0], 0))>::type;
  R acc = R();
  const int n = static_cast<int>(c.size());
  for (int i = 0; i < n; ++i) acc += f(c[i], i);
  return acc;
}

template <typename F>
auto sum(int begin, int end, F&& f) -> typename std::decay<decltype(f(begin))>::type {
  using R = typename std::decay<decltype(f(begin))>::type;
  R acc = R();
  for (int i = begin; i < end; ++i) acc += f(i);
  return acc;
}

}  // namespace futilities
#endif
