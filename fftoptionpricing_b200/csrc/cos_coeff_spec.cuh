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
template <int BASE, int TC, int GMASK, class CP>
__device__ __noinline__ void coeff_det_cold(const double* T, int M, double r, int greek_mask, size_t plane,
                                            unsigned char* C8, unsigned char* rowflag,
                                            const ModelConsts* mcp, double uy, double v, CP crow,
                                            int Kp) {
  CosCoeffArgs a;
  a.T = T;  a.M = M;  a.r = r;  a.greek_mask = greek_mask;  a.plane = plane;  a.npow = nullptr;  a.dt = 0.0;
  a.C8 = C8;  a.rowflag = rowflag;  a.Kp = Kp;
  // (the two-set kernel, GMASK = 9, reads its theta scale from vk[K]: 2^45 / (4 + |r|) as the host sets it)
  const double theta_scale = 35184372088832.0 / (4.0 + fabs(r));
  a.vk = &theta_scale;  a.K = 0;
  const ModelConsts mc = *mcp;
  const cd u = mk(0.0, uy);
  CfPre pre;
  cf_prepare(BASE, TC, mc, u, r, pre);
  coeff_maturities_g<CF_MODE_DET, 1, false, GMASK, false, CP>(a, mc, pre, u, v, crow, Kp);
}

// Q8: emit the coefficients as fixed-point digit planes for the int8 contraction (price only)
// KPC > 0: the padded u count a.Kp as a compile-time constant (the digit-plane stores then use constant offsets)
template <int BASE, int TC, int ILP, bool GRID, int GMASK, int NT, int MINB, bool FAST = true, bool Q8 = false,
          int KPC = 0>
__global__ void __launch_bounds__(NT, MINB) cos_coeff_kernel_s(const CosCoeffArgs a) {
  static_assert(!Q8 || GMASK == 1 || GMASK == 8 || GMASK == 9, "the quantised output carries the price and / or the theta planes");
  using CP = typename std::conditional<Q8, Q8Row, double2*>::type;
  __shared__ ModelConsts consts[COEFF_MAX_SPB];
  const int tid = threadIdx.x;
  const int K = a.K, M = a.M, Kp = KPC > 0 ? KPC : a.Kp;
  const int ngroups = (a.P + a.spb - 1) / a.spb;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int p0 = grp * a.spb;
    const int nsets = min(a.spb, a.P - p0);
    __syncthreads();
    if (tid < nsets) make_consts(BASE, TC, a.params + (size_t)(p0 + tid) * a.np, consts[tid]);
    __syncthreads();
    for (int idx = tid; idx < nsets * Kp; idx += NT) {
      const int s = idx / Kp, k = idx - s * Kp;
      CP crow;
      if constexpr (Q8) {
        crow.p = reinterpret_cast<unsigned short*>(a.C8) + (size_t)(p0 + s) * M * Q8_SLICES * Kp + k;
        if (k >= K) {
          for (int m = 0; m < M * Q8_SLICES; ++m) crow.p[(size_t)m * Kp] = 0;
          if (GMASK == 9)
            for (int m = 0; m < M * Q8_SLICES; ++m) crow.p[a.plane / 2 + (size_t)m * Kp] = 0;
          continue;
        }
      } else {
        crow = reinterpret_cast<double2*>(a.C) + (size_t)(p0 + s) * M * Kp + k;
        if (k >= K) {  // zero padding so the contraction can run full stages
          double2* cg = crow;
          for (int g = 0; g < 4; ++g) {
            if (!((GMASK >= 0 ? GMASK : a.greek_mask) >> g & 1)) continue;
            for (int m = 0; m < M; ++m) cg[(size_t)m * Kp] = make_double2(0.0, 0.0);
            cg += a.plane / 2;
          }
          continue;
        }
      }
      const ModelConsts& mc = consts[s];
      const cd u = mk(0.0, __ldg(a.uk + k));
      const double v = __ldg(a.vk + k);
      if (TC == TC_NONE) {
        CfPre pre;
        cf_prepare(BASE, TC, mc, u, a.r, pre);
        coeff_maturities_g<CF_MODE_LEVY, ILP, false, GMASK, false, CP>(a, mc, pre, u, v, crow, Kp);
      } else if (mc.adaV == 0.0) {
        coeff_det_cold<BASE, TC, GMASK, CP>(a.T, M, a.r, a.greek_mask, a.plane, a.C8, a.rowflag, &consts[s], u.y, v,
                                            crow, Kp);
      } else {
        CfPre pre;
        cf_prepare(BASE, TC, mc, u, a.r, pre);
        coeff_maturities_g<CF_MODE_CIR, ILP, GRID, GMASK, FAST, CP>(a, mc, pre, u, v, crow, Kp);
      }
    }
  }
}

}  // namespace fop
