// Specialised COS coefficient kernels, Levy base VG (cos_coeff_launch.cuh).
#include "cos_coeff_launch.cuh"
namespace fop {
CoeffKernel coeff_pick_vg(int tc, bool grid, int gsel, int shape, int* nt) {
  return coeff_pick_base<BASE_VG>(tc, grid, gsel, shape, nt);
}
}  // namespace fop
