// Drop-in for the reference's use of cuckoo::optimize / swarm_utils (generateCharts.cpp:25,115-118;
// library danielhstahl/cuckoo_search, un-vendored): the search runs on the device over the closure
// returned by optioncal::getObjFn_arr (fop_calibrate_cf).  Same call shape:
//   auto results = cuckoo::optimize(objFn, constraints, nestSize, numSims, tol, seed);
//   auto params  = std::get<swarm_utils::optparms>(results);
//   auto fnval   = std::get<swarm_utils::fnval>(results);
// Extra (defaulted) argument: `restarts` independent searches run concurrently, one CTA each; the
// best is returned.
#ifndef FFTOP_CUCKOO_H
#define FFTOP_CUCKOO_H
#include <tuple>
#include <vector>

#include "OptionCalibration.h"

namespace swarm_utils {
template <typename T>
struct upper_lower {
  T lower, upper;
  upper_lower(T l, T u) : lower(l), upper(u) {}
};
constexpr int optparms = 0;
constexpr int fnval = 1;
}  // namespace swarm_utils

namespace cuckoo {
inline std::tuple<std::vector<double>, double> optimize(
    const optioncal::ObjFn& objFn, const std::vector<swarm_utils::upper_lower<double>>& constraints,
    int nestSize, int numSims, double tol, unsigned long long seed, int restarts = 1) {
  auto& e = fftop::Engine::instance();
  const int np = (int)constraints.size();
  std::vector<double> lo(np), hi(np), prm((size_t)restarts * np), fv(restarts);
  for (int i = 0; i < np; ++i) { lo[i] = constraints[i].lower; hi[i] = constraints[i].upper; }
  int best = 0;
  e.check(fop_calibrate_cf(e.get(), objFn.model, lo.data(), hi.data(), objFn.uArray.data(),
                           reinterpret_cast<const double*>(objFn.phiHat.data()),
                           (int)objFn.uArray.size(), objFn.r, objFn.T, nestSize, numSims, tol, seed,
                           restarts, prm.data(), fv.data(), nullptr, &best));
  return std::make_tuple(std::vector<double>(prm.begin() + (size_t)best * np, prm.begin() + (size_t)(best + 1) * np),
                         fv[best]);
}
}  // namespace cuckoo
#endif
