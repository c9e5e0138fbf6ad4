// Shared pieces of the Fang-Oosterlee COS path (replaces OptionPricing.h:325-346 +
// fangoost::computeExpectationVectorLevy):
//
//   val[row][j] = sum_k  Re(D_k) V_k cos(u_k (x_j - a))  -  Im(D_k) V_k sin(u_k (x_j - a))
//               = sum_kk C[row][kk] * Bas[kk][j],      kk = 0 .. 2 numU - 1
//   row = (parameter set p, maturity m),  D_k = enhCF(phi_{p,m}(i u_k), i u_k) cp  (k = 0 halved)
//
// The basis Bas[2 numU][J] depends only on the strike list, so it is shared by every row: the
// work is a tall-skinny real fp64 GEMM.  Kernels: cos_split_kernels.cuh.
//
// History (profiles/): a first version fused CF + contraction in one CTA; its 244-register
// accumulator tile capped occupancy at 8 warps/SM and left the fp64 pipe 33 % busy
// (r1_cos_fused_v0_ncu_summary.json), so the path was split in two kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cf_models.cuh"

namespace fop {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// basis[2k][j] = cos(u_k (x_j - xMin)), basis[2k+1][j] = -sin(u_k (x_j - xMin)); 0 for j >= J
static __global__ void cos_basis_kernel(const double* __restrict__ uk, const double* __restrict__ x,
                                 int K, int J, int Jp, double* __restrict__ basis) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Jp) return;
  const int k = idx / Jp, j = idx - k * Jp;
  double c = 0.0, s = 0.0;
  if (j < J) sincos(uk[k] * (x[j] - x[0]), &s, &c);
  basis[(size_t)(2 * k) * Jp + j] = c;
  basis[(size_t)(2 * k + 1) * Jp + j] = -s;
}

}  // namespace fop
