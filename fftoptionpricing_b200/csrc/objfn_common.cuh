// Warp-level evaluation of the calibration objective of optioncal::getObjFn_arr
// (OptionCalibration.h:185-201), shared by fop_objfn_cf (one pass over P parameter vectors) and
// fop_calibrate_cf (the on-device population search): identical code => identical bits.
//   loss = (1/m) sum_l ( isnan(c) ? 5000 : |phiHat_l - c|^2 ),   c = logCF(1 + i u_l)
#pragma once
#include "cf_models.cuh"

namespace fop {

constexpr double kLargeNumber = 5000.0;  // OptionCalibration.h:185

// All 32 lanes of a warp call this with the same arguments; `mc` and `terms` ([m] doubles) are the
// warp's scratch in shared memory.  Lanes stride over u_l; the m terms are summed in index order
// by lane 0 (the order of futilities::sum in the reference -> deterministic) and broadcast.
__device__ __forceinline__ double warp_objfn_cf(int base, int tc, const double* __restrict__ theta,
                                                int m, const double* __restrict__ u,
                                                const double* __restrict__ phiHat, double r, double T,
                                                ModelConsts& mc, double* terms) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) make_consts(base, tc, theta, mc);
  __syncwarp();
  for (int l = lane; l < m; l += 32) {
    const cd uu = mk(1.0, __ldg(u + l));
    CfPre pre;
    cf_prepare(base, tc, mc, uu, r, pre);
    const cd c = cf_log_at(tc, mc, pre, T);
    const double dr = __ldg(phiHat + 2 * l) - c.x, di = __ldg(phiHat + 2 * l + 1) - c.y;
    terms[l] = (isnan(c.x) || isnan(c.y)) ? kLargeNumber : dr * dr + di * di;
  }
  __syncwarp();
  double acc = 0.0;
  if (lane == 0) {
    for (int l = 0; l < m; ++l) acc += terms[l];
    acc = acc / (double)m;
  }
  acc = __shfl_sync(0xffffffffu, acc, 0);
  __syncwarp();
  return acc;
}

}  // namespace fop
