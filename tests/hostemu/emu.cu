// TEST INFRASTRUCTURE ONLY -- CPU emulation of the device code of the COS coefficient kernel.
//
// The characteristic-function layer of the product (csrc/fmath.cuh, cplx.cuh, cf_models.cuh,
// cos_coeff.cuh) is __host__ __device__ source; this file compiles THE SAME SOURCE for the host and
// runs it over whole parameter sweeps, so that branch-cut / special-case behaviour of the
// hand-written elementary functions can be compared with the oracle on a machine without a GPU
// (tests/test_hostemu.py, scripts/sweep_hostemu.py).  It is not a fallback: nothing in the
// package loads it, and the rounding differs from the device in the last bits (host FMA
// contraction), which is irrelevant for what it checks.  The contraction is a plain loop.
//
// Build: nvcc -O2 -std=c++17 -Xcompiler -fPIC,-fopenmp,-mfma -shared -I fftoptionpricing_b200/csrc
#include <cmath>
#include <cstring>
#include <vector>

#include "cos_coeff.cuh"
#include "cos_host_tables.h"

using namespace fop;

extern "C" int emu_cos(int base, int tc, const double* params, int P, const double* K_desc, int J,
                       double S0, double r, const double* T, int M, int numU, int kind,
                       int use_tgrid, int fast, double* out) {
  const int np = num_params(base, tc);
  if (np == 0 || J < 2 || M < 1 || numU < 1 || kind < 0 || kind > 7) return 1;
  std::vector<double> Th(T, T + M), x(J), uk(numU), vk(numU);
  for (int j = 0; j < J; ++j) x[j] = std::log(S0 / K_desc[j]);
  const double xMin = x[0], xMax = x[J - 1];
  const double du = M_PI / (xMax - xMin), cp = 2.0 / (xMax - xMin);
  for (int k = 0; k < numU; ++k) {
    uk[k] = du * k;
    vk[k] = vk_put(xMin, uk[k], k) * cp * (k == 0 ? 0.5 : 1.0);
  }
  double dt = 0.0;
  std::vector<int> npow, steps;
  if (use_tgrid && maturity_grid(Th, &dt, &npow)) steps = maturity_steps(npow);
  if (steps.empty()) dt = 0.0;
  std::vector<double> bc((size_t)numU * J), bs((size_t)numU * J);
  for (int k = 0; k < numU; ++k)
    for (int j = 0; j < J; ++j) {
      bc[(size_t)k * J + j] = std::cos(uk[k] * (x[j] - x[0]));
      bs[(size_t)k * J + j] = -std::sin(uk[k] * (x[j] - x[0]));
    }
  const CosEpi epi = cos_epi_consts(kind, S0, r);
  const int Kp = numU;
#pragma omp parallel
  {
    std::vector<double2> C((size_t)M * Kp);
#pragma omp for schedule(dynamic, 16)
    for (int p = 0; p < P; ++p) {
      CosCoeffArgs a;
      a.base = base;  a.tc = tc;  a.greek_mask = 1 << (kind >> 1);  a.plane = 0;
      a.P = 1;  a.M = M;  a.K = numU;  a.Kp = Kp;  a.np = np;  a.spb = 1;  a.r = r;
      a.params = params + (size_t)p * np;  a.T = Th.data();
      a.npow = dt > 0.0 ? steps.data() : nullptr;  a.dt = dt;
      a.uk = uk.data();  a.vk = vk.data();  a.C = reinterpret_cast<double*>(C.data());
      a.C8 = nullptr;  a.rowflag = nullptr;
      ModelConsts mc;
      make_consts(base, tc, a.params, mc);
      const int mode = cf_mode(tc, mc);
      for (int k = 0; k < numU; ++k) {
        const cd u = mk(0.0, uk[k]);
        CfPre pre;
        cf_prepare(base, tc, mc, u, r, pre);
        double2* crow = C.data() + k;
        if (mode == CF_MODE_CIR && fast) {  // the specialised kernels' path (cos_coeff_spec.cuh)
          const int gm = a.greek_mask;
          if (a.npow) {
            if (gm == 1) coeff_maturities_g<CF_MODE_CIR, 2, true, 1, true>(a, mc, pre, u, vk[k], crow, Kp);
            else coeff_maturities_g<CF_MODE_CIR, 2, true, -1, true>(a, mc, pre, u, vk[k], crow, Kp);
          } else {
            if (gm == 1) coeff_maturities_g<CF_MODE_CIR, 2, false, 1, true>(a, mc, pre, u, vk[k], crow, Kp);
            else coeff_maturities_g<CF_MODE_CIR, 2, false, -1, true>(a, mc, pre, u, vk[k], crow, Kp);
          }
        } else if (mode == CF_MODE_CIR) coeff_maturities<CF_MODE_CIR, 2>(a, mc, pre, u, vk[k], crow, Kp);
        else if (mode == CF_MODE_LEVY) coeff_maturities<CF_MODE_LEVY, 2>(a, mc, pre, u, vk[k], crow, Kp);
        else coeff_maturities<CF_MODE_DET, 1>(a, mc, pre, u, vk[k], crow, Kp);
      }
      for (int m = 0; m < M; ++m) {
        const double rs = std::exp(-r * Th[m]) * epi.inv;
        double* o = out + ((size_t)p * M + m) * J;
        for (int j = 0; j < J; ++j) o[j] = 0.0;
        for (int k = 0; k < numU; ++k) {
          const double2 c = C[(size_t)m * Kp + k];
          const double* pc = &bc[(size_t)k * J];
          const double* ps = &bs[(size_t)k * J];
          for (int j = 0; j < J; ++j) o[j] += c.x * pc[j] + c.y * ps[j];
        }
        for (int j = 0; j < J; ++j) o[j] = (o[j] - epi.sub) * (rs * K_desc[j]) + epi.add;
      }
    }
  }
  return 0;
}

// log CF at (params, u) through the device code (prepare + per-maturity step), for point probes
extern "C" int emu_logcf(int base, int tc, const double* params, double r, double T, double u_re,
                         double u_im, double* out2) {
  if (num_params(base, tc) == 0) return 1;
  ModelConsts mc;
  make_consts(base, tc, params, mc);
  CfPre pre;
  cf_prepare(base, tc, mc, mk(u_re, u_im), r, pre);
  const cd l = cf_log_at(tc, mc, pre, T);
  out2[0] = l.x;
  out2[1] = l.y;
  return 0;
}

// Digit-plane store of the int8 contraction (cos_coeff.cuh: coeff_store(Q8Row), the very code the
// quantising coefficient kernels run): n complex values -> [n][6][Kp = 1] byte pairs + row flags, and
// the integers the six signed digits reassemble to.
extern "C" int emu_q8_store(const double* re, const double* im, int n, unsigned char* planes,
                            unsigned char* rowflag, long long* q_re, long long* q_im) {
  CosCoeffArgs a{};
  a.Kp = 1;
  a.C8 = planes;
  a.rowflag = rowflag;
  for (int i = 0; i < n; ++i) {
    Q8Row row;
    row.p = reinterpret_cast<unsigned short*>(planes) + (size_t)i * Q8_SLICES;
    coeff_store(a, row, 0, 1, re[i], im[i]);
    long long qr = 0, qi = 0;
    for (int s = Q8_SLICES - 1; s >= 0; --s) {
      qr = qr * 256 + (signed char)planes[((size_t)i * Q8_SLICES + s) * 2];
      qi = qi * 256 + (signed char)planes[((size_t)i * Q8_SLICES + s) * 2 + 1];
    }
    q_re[i] = qr;
    q_im[i] = qi;
  }
  return 0;
}
