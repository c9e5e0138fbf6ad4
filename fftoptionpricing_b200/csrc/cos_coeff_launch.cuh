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
  // shape >= 100: quantised output (digit planes for the int8 contraction; one transform)
  if constexpr (GMASK == 8 || GMASK == 9) {
    if (shape >= 100) {
      *nt = 256;
      return cos_coeff_kernel_s<BASE, TC, 2, GRID, GMASK, 256, 2, true, true>;
    }
    return nullptr;
  }
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

// gsel: -1 = run-time transform mask (fp64 output), 1 = price only, 8 = theta only, 9 = price and theta planes (quantised output only;
// delta and gamma are contracted from the price planes)
template <int BASE, int TC, bool GRID>
CoeffKernel coeff_pick_mask(int gsel, int shape, int* nt) {
  if (gsel == 1) return coeff_pick_shape<BASE, TC, GRID, 1>(shape, nt);
  if (gsel == 8) return coeff_pick_shape<BASE, TC, GRID, 8>(shape, nt);
  if (gsel == 9) return coeff_pick_shape<BASE, TC, GRID, 9>(shape, nt);
  return coeff_pick_shape<BASE, TC, GRID, -1>(shape, nt);
}

template <int BASE, int TC>
CoeffKernel coeff_pick_tc(bool grid, int gsel, int shape, int* nt) {
  if constexpr (TC == TC_NONE) {
    return coeff_pick_mask<BASE, TC, false>(gsel, shape, nt);
  } else {
    if (grid) return coeff_pick_mask<BASE, TC, true>(gsel, shape, nt);
    return coeff_pick_mask<BASE, TC, false>(gsel, shape, nt);
  }
}

template <int BASE>
CoeffKernel coeff_pick_base(int tc, bool grid, int gsel, int shape, int* nt) {
  if (tc == TC_NONE) return coeff_pick_tc<BASE, TC_NONE>(grid, gsel, shape, nt);
  if (tc == TC_CIR) return coeff_pick_tc<BASE, TC_CIR>(grid, gsel, shape, nt);
  return coeff_pick_tc<BASE, TC_CIR_DIFFUSION>(grid, gsel, shape, nt);
}

CoeffKernel coeff_pick_gauss(int tc, bool grid, int gsel, int shape, int* nt);
CoeffKernel coeff_pick_merton(int tc, bool grid, int gsel, int shape, int* nt);
CoeffKernel coeff_pick_cgmy(int tc, bool grid, int gsel, int shape, int* nt);
CoeffKernel coeff_pick_vg(int tc, bool grid, int gsel, int shape, int* nt);

}  // namespace fop
