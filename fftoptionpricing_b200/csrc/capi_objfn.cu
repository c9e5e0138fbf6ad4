// C ABI: fop_objfn_cf -- the closure of optioncal::getObjFn_arr (OptionCalibration.h:185-201)
// evaluated for P parameter vectors at once:
//   loss_p = (1/m) sum_l ( isnan(c) ? 5000 : |phiHat_l - c|^2 ),   c = logCF_p(1 + i u_l)
#include <algorithm>
#include <cmath>

#include "handle.cuh"
#include "objfn_common.cuh"

namespace {
using namespace fop;

constexpr int OBJ_WARPS = 8;

struct ObjArgs {
  int base, tc, np, P, m;
  double r, T;
  const double* params;
  const double* u;       // [m]
  const double* phiHat;  // [m][2]
  double* loss;          // [P]
};

// one warp per parameter set, lanes over u_l; the m terms are summed in index order by lane 0
// (same order as futilities::sum in the reference -> deterministic)
__global__ void __launch_bounds__(32 * OBJ_WARPS) objfn_cf_kernel(const ObjArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ModelConsts* consts = reinterpret_cast<ModelConsts*>(smem_raw);
  double* terms = reinterpret_cast<double*>(smem_raw + sizeof(ModelConsts) * OBJ_WARPS) + (size_t)warp * a.m;
  for (int p = blockIdx.x * OBJ_WARPS + warp; p < a.P; p += gridDim.x * OBJ_WARPS) {
    const double v = warp_objfn_cf(a.base, a.tc, a.params + (size_t)p * a.np, a.m, a.u, a.phiHat, a.r,
                                   a.T, consts[warp], terms);
    if (lane == 0) a.loss[p] = v;
  }
}

}  // namespace

extern "C" int fop_objfn_cf(fop_handle_t h, fop_model_t model, const double* params, int P,
                            const double* u, const double* phiHat, int m, double r, double T,
                            double* loss) {
  if (!h) return FOP_INVALID_ARG;
  const int np = num_params(model.base, model.time_change);
  if (np == 0) return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if (!params || !u || !phiHat || !loss) return fail(h, FOP_INVALID_ARG, "null pointer");
  if (P < 0 || m < 1 || m > 2048) return fail(h, FOP_INVALID_ARG, "fop_objfn_cf: need 1 <= m <= 2048");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const bool loss_host = !is_device_ptr(loss);
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(params, (size_t)P * np * 8) + stage_bytes(u, (size_t)m * 8) +
                     stage_bytes(phiHat, (size_t)m * 16) + (loss_host ? Arena::pad((size_t)P * 8) : 0)));
  ObjArgs a;
  a.base = model.base;  a.tc = model.time_change;  a.np = np;  a.P = P;  a.m = m;  a.r = r;  a.T = T;
  FOP_TRY(stage_in(h, ar, params, (size_t)P * np, &a.params));
  FOP_TRY(stage_in(h, ar, u, (size_t)m, &a.u));
  FOP_TRY(stage_in(h, ar, phiHat, (size_t)2 * m, &a.phiHat));
  a.loss = loss_host ? ar.alloc<double>(P) : loss;
  const size_t smem = sizeof(ModelConsts) * OBJ_WARPS + (size_t)OBJ_WARPS * m * 8;
  FOP_CUDA(h, cudaFuncSetAttribute(objfn_cf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min((P + OBJ_WARPS - 1) / OBJ_WARPS, h->sm_count * 8);
  {
    ProfScope ps(h, "objfn_cf_kernel");
    objfn_cf_kernel<<<grid, 32 * OBJ_WARPS, smem, h->stream>>>(a);
  }
  FOP_LAUNCH_CHECK(h);
  if (loss_host) {
    FOP_CUDA(h, cudaMemcpyAsync(loss, a.loss, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}
