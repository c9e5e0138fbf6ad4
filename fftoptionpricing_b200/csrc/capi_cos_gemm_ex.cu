// Instantiations of the COS contraction kernel with 1 or 2 remainder columns (J = 8 NB + EX).
#include "cos_gemm_variants.h"

namespace fop {
namespace {
const GemmExVariant kGemmEx[] = {{4, 1, cos_gemm_dmma_kernel<4, 4, 1>}, {4, 2, cos_gemm_dmma_kernel<4, 4, 2>},
                                 {8, 1, cos_gemm_dmma_kernel<8, 4, 1>}, {8, 2, cos_gemm_dmma_kernel<8, 4, 2>}};
}
const GemmExVariant* gemm_ex_variants(int* count) {
  *count = (int)(sizeof(kGemmEx) / sizeof(kGemmEx[0]));
  return kGemmEx;
}
}  // namespace fop
