// Instantiations of the TMA / mbarrier COS contraction kernel (cos_gemm_tma.cuh).
#include "cos_gemm_tma.cuh"

namespace fop {
namespace {
#define FOP_TMA_V(NB, EX) {NB, EX, cos_gemm_tma_kernel<NB, EX, 4>, cos_gemm_tma_kernel<NB, EX, 2>, nullptr}
const GemmTmaVariant kGemmTma[] = {
    FOP_TMA_V(2, 0), FOP_TMA_V(4, 0), FOP_TMA_V(6, 0),
    {8, 0, cos_gemm_tma_kernel<8, 0, 4>, cos_gemm_tma_kernel<8, 0, 2>, cos_gemm_tma_kernel<8, 0, 1>},
    FOP_TMA_V(9, 0), FOP_TMA_V(4, 1), FOP_TMA_V(4, 2), FOP_TMA_V(8, 1),
    {8, 2, cos_gemm_tma_kernel<8, 2, 4>, cos_gemm_tma_kernel<8, 2, 2>, cos_gemm_tma_kernel<8, 2, 1>}};
#undef FOP_TMA_V
}
const GemmTmaVariant* gemm_tma_variants(int* count) {
  *count = (int)(sizeof(kGemmTma) / sizeof(kGemmTma[0]));
  return kGemmTma;
}
}  // namespace fop
