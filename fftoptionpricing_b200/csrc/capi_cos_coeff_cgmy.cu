// Specialised COS coefficient kernels, Levy base CGMY (cos_coeff_launch.cuh).
#include "cos_coeff_launch.cuh"
namespace fop {
CoeffKernel coeff_pick_cgmy(int tc, bool grid, int gsel, int shape, int* nt) {
  return coeff_pick_base<BASE_CGMY>(tc, grid, gsel, shape, nt);
}
}  // namespace fop
