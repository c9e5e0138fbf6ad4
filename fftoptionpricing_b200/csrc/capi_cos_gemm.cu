// Instantiations of the COS contraction kernel (cos_split_kernels.cuh), strike blocks of 8 NB columns.
#include "cos_gemm_variants.h"

namespace fop {
namespace {
const GemmVariant kGemm[] = {{2, cos_gemm_dmma_kernel<2, 4>, cos_gemm_dmma_kernel<2, 2>},
                             {4, cos_gemm_dmma_kernel<4, 4>, cos_gemm_dmma_kernel<4, 2>},
                             {6, cos_gemm_dmma_kernel<6, 4>, cos_gemm_dmma_kernel<6, 2>},
                             {8, cos_gemm_dmma_kernel<8, 4>, cos_gemm_dmma_kernel<8, 2>},
                             {9, cos_gemm_dmma_kernel<9, 4>, cos_gemm_dmma_kernel<9, 2>}};
}
const GemmVariant* gemm_variants(int* count) {
  *count = (int)(sizeof(kGemm) / sizeof(kGemm[0]));
  return kGemm;
}
}  // namespace fop
