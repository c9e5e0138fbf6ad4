// Instantiation table of the model-specialised COS coefficient kernel for ONE Levy base (included
// by capi_cos_coeff_{gauss,merton,cgmy}.cu, one translation unit per base so that they compile in
// parallel).
#pragma once
#include "cos_coeff_spec.cuh"

namespace fop {

typedef void (*CoeffKernel)(const CosCoeffArgs);

// CTA shapes of the calibration-sweep model (GAUSS + CIR, grid maturities, price only), selectable
// with FOP_COEFF_SHAPE for A/B runs; every other combination uses 256 threads x 2 CTAs/SM, ILP 2.
struct CoeffShape { int ilp, nt, minb; };
constexpr CoeffShape kCoeffShapes[] = {{2, 256, 2}, {2, 128, 5}, {4, 128, 3}};
constexpr int kNumCoeffShapes = sizeof(kCoeffShapes) / sizeof(kCoeffShapes[0]);

template <int BASE, int TC, bool GRID, int GMASK>
CoeffKernel coeff_pick_shape(int shape, int* nt) {
  // shape >= 100: quantised output (digit planes for the int8 contraction; price only)
  if constexpr (GMASK == 1) {
    if (shape >= 100) {
      if constexpr (BASE == BASE_GAUSS && TC == TC_CIR && GRID) {
        if (shape == 101) { *nt = 128; return cos_coeff_kernel_s<BASE, TC, 2, GRID, GMASK, 128, 5, true, true>; }
        if (shape == 102) { *nt = 128; return cos_coeff_kernel_s<BASE, TC, 2, GRID, GMASK, 128, 5, true, true, 256>; }
      }
      *nt = 256;
      return cos_coeff_kernel_s<BASE, TC, 2, GRID, GMASK, 256, 2, true, true>;
    }
  }
  if (shape >= 100) return nullptr;
  if constexpr (BASE == BASE_GAUSS && TC == TC_CIR && GRID && GMASK == 1) {
    if (shape == 1) { *nt = 128; return cos_coeff_kernel_s<BASE, TC, 2, GRID, GMASK, 128, 5>; }
    if (shape == 2) { *nt = 128; return cos_coeff_kernel_s<BASE, TC, 4, GRID, GMASK, 128, 3>; }
  }
  *nt = 256;
  return cos_coeff_kernel_s<BASE, TC, 2, GRID, GMASK, 256, 2>;
}

template <int BASE, int TC>
CoeffKernel coeff_pick_tc(bool grid, bool price_only, int shape, int* nt) {
  if constexpr (TC == TC_NONE) {
    return price_only ? coeff_pick_shape<BASE, TC, false, 1>(shape, nt) : coeff_pick_shape<BASE, TC, false, -1>(shape, nt);
  } else {
    if (grid) return price_only ? coeff_pick_shape<BASE, TC, true, 1>(shape, nt) : coeff_pick_shape<BASE, TC, true, -1>(shape, nt);
    return price_only ? coeff_pick_shape<BASE, TC, false, 1>(shape, nt) : coeff_pick_shape<BASE, TC, false, -1>(shape, nt);
  }
}

template <int BASE>
CoeffKernel coeff_pick_base(int tc, bool grid, bool price_only, int shape, int* nt) {
  if (tc == TC_NONE) return coeff_pick_tc<BASE, TC_NONE>(grid, price_only, shape, nt);
  if (tc == TC_CIR) return coeff_pick_tc<BASE, TC_CIR>(grid, price_only, shape, nt);
  return coeff_pick_tc<BASE, TC_CIR_DIFFUSION>(grid, price_only, shape, nt);
}

CoeffKernel coeff_pick_gauss(int tc, bool grid, bool price_only, int shape, int* nt);
CoeffKernel coeff_pick_merton(int tc, bool grid, bool price_only, int shape, int* nt);
CoeffKernel coeff_pick_cgmy(int tc, bool grid, bool price_only, int shape, int* nt);
CoeffKernel coeff_pick_vg(int tc, bool grid, bool price_only, int shape, int* nt);

}  // namespace fop
