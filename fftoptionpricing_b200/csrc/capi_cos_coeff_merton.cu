// Specialised COS coefficient kernels, Levy base MERTON (cos_coeff_launch.cuh).
#include "cos_coeff_launch.cuh"
namespace fop {
CoeffKernel coeff_pick_merton(int tc, bool grid, int gsel, int shape, int* nt) {
  return coeff_pick_base<BASE_MERTON>(tc, grid, gsel, shape, nt);
}
}  // namespace fop
