// Tuning probe (not part of the public header): the int8 tensor rate of tcgen05.mma.kind::i8 with both
// operands resident in shared memory -- the denominator bench.py reports for the digit-plane
// contraction (cos_gemm_i8.cuh) next to NVIDIA's nominal figure.  One CTA per SM, one thread issues
// M 128 x N 256 x K 32 instructions back to back into two alternating halves of TMEM.
#include "cos_gemm_i8.cuh"
#include "handle.cuh"

namespace {
using namespace fop;

__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_done;
  __shared__ uint32_t s_tmem;
  unsigned char* base = smem_dyn + ((1024u - (tm_smem_u32(smem_dyn) & 1023u)) & 1023u);
  // A: 128 rows x 128 B, B: 256 rows x 128 B (K-major, 128-byte swizzle); the values do not matter
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(base)[i] = 0x01ff02feu * (uint32_t)(i + 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    tm_mbar_init(&s_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(&s_tmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t sa = tm_smem_u32(base), sb = sa + 128 * 128;
    const uint32_t idesc = q8_idesc(256);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        q8_mma(tmem + (it & 1) * 256, q8_smem_desc(sa + ks * 32), q8_smem_desc(sb + ks * 32), idesc, (it > 1 || ks > 0) ? 1u : 0u);
    }
    tc_commit(&s_done);
    q8_wait(&s_done, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int fop_microbench_int8(fop_handle_t h, double* pops_out) {
  if (!h || !pops_out) return FOP_INVALID_ARG;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const int iters = 1 << 14, grid = h->sm_count;
  const size_t smem = (128 + 256) * 128 + 1024;
  FOP_CUDA(h, cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  FOP_CUDA(h, cudaEventCreate(&e0));
  FOP_CUDA(h, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    FOP_CUDA(h, cudaEventRecord(e0, h->stream));
    i8_peak_kernel<<<grid, 128, smem, h->stream>>>(iters);
    FOP_CUDA(h, cudaEventRecord(e1, h->stream));
    FOP_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    FOP_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  FOP_LAUNCH_CHECK(h);
  h->launches += 4;
  // per CTA per iteration: 4 instructions of 128 x 256 x 32 multiply-adds
  *pops_out = 2.0 * 4 * 128.0 * 256.0 * 32.0 * iters * grid / (best * 1e-3) / 1e15;
  return FOP_OK;
}
