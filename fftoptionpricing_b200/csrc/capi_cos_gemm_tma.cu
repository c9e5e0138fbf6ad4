// Instantiations of the TMA / mbarrier COS contraction kernel (cos_gemm_tma.cuh).
#include "cos_gemm_tma.cuh"

namespace fop {
namespace {
const GemmTmaVariant kGemmTma[] = {
    {2, 0, cos_gemm_tma_kernel<2, 0>}, {4, 0, cos_gemm_tma_kernel<4, 0>}, {6, 0, cos_gemm_tma_kernel<6, 0>},
    {8, 0, cos_gemm_tma_kernel<8, 0>}, {9, 0, cos_gemm_tma_kernel<9, 0>}, {4, 1, cos_gemm_tma_kernel<4, 1>},
    {4, 2, cos_gemm_tma_kernel<4, 2>}, {8, 1, cos_gemm_tma_kernel<8, 1>}, {8, 2, cos_gemm_tma_kernel<8, 2>}};
}
const GemmTmaVariant* gemm_tma_variants(int* count) {
  *count = (int)(sizeof(kGemmTma) / sizeof(kGemmTma[0]));
  return kGemmTma;
}
}  // namespace fop
