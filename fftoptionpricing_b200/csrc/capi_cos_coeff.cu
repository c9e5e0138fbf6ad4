// Launcher of the model-specialised COS coefficient kernels (cos_coeff_spec.cuh; instantiated per
// Levy base in capi_cos_coeff_{gauss,merton,cgmy}.cu).
#include <algorithm>
#include <cstdlib>

#include "cos_coeff_launch.cuh"
#include "handle.cuh"

namespace fop {

// Launches the specialised coefficient kernel for `a_in` on the handle's stream; returns false when
// no specialisation exists (the caller then uses the generic kernel).
bool cos_coeff_launch_spec(fop_handle_t h, const CosCoeffArgs& a_in, int grid_blocks_per_sm) {
  CosCoeffArgs a = a_in;
  int shape = 1;  // 128 threads x 5 CTAs/SM: measured best on B200 (profiles/r2_coeff_variants.json)
  if (const char* e = std::getenv("FOP_COEFF_SHAPE")) shape = std::max(0, std::min(atoi(e), kNumCoeffShapes - 1));
  if (a.M < kCoeffShapes[shape].ilp) shape = 1;
  // quantised output (cos_gemm_i8.cuh); 102: a.Kp == 256 compiled in (GAUSS + CIR on a maturity grid)
  if (a.C8) shape = shape == 1 ? ((a.Kp == 256 && a.base == BASE_GAUSS && a.tc == TC_CIR && a.npow) ? 102 : 101) : 100;
  if (a.C8 && a.greek_mask != 1) shape = 100;
  const bool grid = a.npow != nullptr;
  // one compile-time transform: price (both outputs; its planes also serve delta and gamma), or theta
  // with the quantised output
  const int gsel = a.greek_mask == 1 ? 1 : (a.C8 && (a.greek_mask == 8 || a.greek_mask == 9)) ? a.greek_mask : -1;
  if (a.C8 && gsel < 0) return false;
  int nt = 256;
  CoeffKernel k = nullptr;
  if (a.base == BASE_GAUSS) k = coeff_pick_gauss(a.tc, grid, gsel, shape, &nt);
  else if (a.base == BASE_MERTON) k = coeff_pick_merton(a.tc, grid, gsel, shape, &nt);
  else if (a.base == BASE_CGMY) k = coeff_pick_cgmy(a.tc, grid, gsel, shape, &nt);
  else if (a.base == BASE_VG) k = coeff_pick_vg(a.tc, grid, gsel, shape, &nt);
  if (!k) return false;
  // sets per CTA: four (set, u) pairs per thread, whole sets only
  a.spb = std::max(1, std::min(COEFF_MAX_SPB, 4 * nt / a.Kp));
  const int ngroups = (a.P + a.spb - 1) / a.spb;
  const int cgrid = std::min(ngroups, h->sm_count * grid_blocks_per_sm);
  k<<<cgrid, nt, 0, h->stream>>>(a);
  return true;
}

}  // namespace fop
