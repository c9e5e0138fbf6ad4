// Tables of the instantiated contraction kernels (defined in capi_cos_gemm.cu / capi_cos_gemm_ex.cu).
#pragma once
#include "cos_split_kernels.cuh"

namespace fop {

struct GemmVariant {
  int NB;                              // strike block JB = 8 NB columns (NB n-blocks of the MMA)
  void (*dmma4)(const CosGemmArgs);    // 4 warps x 32 rows: fewest smem wavefronts per MMA
  void (*dmma2)(const CosGemmArgs);    // 8 warps x 16 rows
};
// one strike block + 1 or 2 remainder columns on DFMAs (J = 8 NB + EX): e.g. the 64 quoted + 2 end
// strikes of a calibration surface cost 8 n-blocks instead of 9
struct GemmExVariant {
  int NB, EX;
  void (*kern)(const CosGemmArgs);
};
const GemmVariant* gemm_variants(int* count);
const GemmExVariant* gemm_ex_variants(int* count);

}  // namespace fop
