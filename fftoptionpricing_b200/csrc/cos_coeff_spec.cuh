// Model-specialised COS coefficient kernel (the default for every model; the generic kernel of
// cos_split_kernels.cuh remains for FOP_COEFF_GENERIC=1 A/B runs).
//
// Compile-time: Levy base, time change, maturities on a grid or not, set of emitted transforms,
// maturities advanced together per thread (ILP), CTA shape.  What that buys over the generic kernel:
// the per-maturity loop is one straight-line body (no run-time model / mask / grid tests inside),
// the rare deterministic-clock sets (adaV == 0, test.cpp:1187) run in a cold out-of-line path that
// does not constrain the register allocation of the hot loop, and the static instruction mix of
// the loop is its dynamic mix (scripts/sass_mix.py).
#pragma once
#include "cos_coeff.cuh"

namespace fop {

// cold path: rebuilt from scalars so that the kernel parameter block is not copied to the stack
template <int BASE, int TC, int GMASK>
__device__ __noinline__ void coeff_det_cold(const double* T, int M, double r, int greek_mask, size_t plane,
                                            const ModelConsts* mcp, double uy, double v, double2* crow,
                                            int Kp) {
  CosCoeffArgs a;
  a.T = T;  a.M = M;  a.r = r;  a.greek_mask = greek_mask;  a.plane = plane;  a.npow = nullptr;  a.dt = 0.0;
  const ModelConsts mc = *mcp;
  const cd u = mk(0.0, uy);
  CfPre pre;
  cf_prepare(BASE, TC, mc, u, r, pre);
  coeff_maturities_g<CF_MODE_DET, 1, false, GMASK>(a, mc, pre, u, v, crow, Kp);
}

template <int BASE, int TC, int ILP, bool GRID, int GMASK, int NT, int MINB, bool FAST = true>
__global__ void __launch_bounds__(NT, MINB) cos_coeff_kernel_s(const CosCoeffArgs a) {
  __shared__ ModelConsts consts[COEFF_MAX_SPB];
  const int tid = threadIdx.x;
  const int K = a.K, M = a.M, Kp = a.Kp;
  const int ngroups = (a.P + a.spb - 1) / a.spb;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int p0 = grp * a.spb;
    const int nsets = min(a.spb, a.P - p0);
    __syncthreads();
    if (tid < nsets) make_consts(BASE, TC, a.params + (size_t)(p0 + tid) * a.np, consts[tid]);
    __syncthreads();
    for (int idx = tid; idx < nsets * Kp; idx += NT) {
      const int s = idx / Kp, k = idx - s * Kp;
      double2* crow = reinterpret_cast<double2*>(a.C) + (size_t)(p0 + s) * M * Kp + k;
      if (k >= K) {  // zero padding so the contraction can run full stages
        double2* cg = crow;
        for (int g = 0; g < 4; ++g) {
          if (!((GMASK >= 0 ? GMASK : a.greek_mask) >> g & 1)) continue;
          for (int m = 0; m < M; ++m) cg[(size_t)m * Kp] = make_double2(0.0, 0.0);
          cg += a.plane / 2;
        }
        continue;
      }
      const ModelConsts& mc = consts[s];
      const cd u = mk(0.0, __ldg(a.uk + k));
      const double v = __ldg(a.vk + k);
      if (TC == TC_NONE) {
        CfPre pre;
        cf_prepare(BASE, TC, mc, u, a.r, pre);
        coeff_maturities_g<CF_MODE_LEVY, ILP, false, GMASK>(a, mc, pre, u, v, crow, Kp);
      } else if (mc.adaV == 0.0) {
        coeff_det_cold<BASE, TC, GMASK>(a.T, M, a.r, a.greek_mask, a.plane, &consts[s], u.y, v, crow, Kp);
      } else {
        CfPre pre;
        cf_prepare(BASE, TC, mc, u, a.r, pre);
        coeff_maturities_g<CF_MODE_CIR, ILP, GRID, GMASK, FAST>(a, mc, pre, u, v, crow, Kp);
      }
    }
  }
}

}  // namespace fop
