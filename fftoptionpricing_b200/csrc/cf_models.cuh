// Characteristic-function library (device, fp64).
//
// Restates, as closed-set device functions, the CFs the reference builds from lambdas over the
// un-vendored `chfunctions::` library (call sites test.cpp:54,570,731,887-905,
// 1049-1058,1087-1096,1119; generateCharts.cpp:190-380):
//   gaussLogCF(u,mu,s)          = u mu + s^2 u^2 / 2
//   mertonLogRNCF(u,l,mJ,sJ,r,s)= gaussLogCF(u, r - s^2/2 - l(e^{mJ+sJ^2/2}-1), s) + l(e^{u mJ+sJ^2u^2/2}-1)
//   cgmyLogRNCF(u,C,G,M,Y,r,s)  = gaussLogCF(u, r - s^2/2 - k(1), s) + k(u),
//                                 k(u) = C Gamma(-Y)[(M-u)^Y + (G+u)^Y - M^Y - G^Y]  (0 if C == 0)
//   vgLogRNCF(u,s,th,nu,r)      = u (r + log(1 - th nu - s^2 nu/2)/nu) - log(1 - u th nu - s^2 nu u^2/2)/nu
//                                 (variance gamma: named by BASELINE.json's north_star, absent from the
//                                 reference -> textbook form, parity unpinned; no Brownian part)
//   cirLogMGF(psi,a,kap,sv,t,v0)= -b v0 - c, delta = sqrt(kap^2 + 2 psi sv^2), e = exp(-delta t),
//                                 b = 2 psi (1-e)/(delta+kap+(delta-kap)e),
//                                 c = (a/sv^2)[2 log(1-(delta-kap)(1-e)/(2 delta)) + (delta-kap) t]
//                                 (sv == 0: deterministic clock, test.cpp:1187)
//   time-changed CF             = exp(r T u + cirLogMGF(-psi_0(u), speed, speed - adaV rho u s, adaV, T, v0Hat))
//
// The evaluation is split in two so that each (parameter set, u) pair is worked out once and
// re-used for every maturity: cf_prepare() does everything that does not depend on T (the Levy
// exponent with its complex powers / exponentials, kappa(u), delta, ...), cf_log_at() the rest.
#pragma once
#include "cplx.cuh"

namespace fop {

enum { BASE_GAUSS = 0, BASE_MERTON = 1, BASE_CGMY = 2, BASE_VG = 3 };
enum { TC_NONE = 0, TC_CIR = 1, TC_CIR_DIFFUSION = 2 };

__host__ __device__ __forceinline__ int num_base_params(int base) {
  return base == BASE_GAUSS ? 1 : base == BASE_MERTON ? 4 : base == BASE_CGMY ? 5 : base == BASE_VG ? 3 : 0;
}
__host__ __device__ __forceinline__ int num_params(int base, int tc) {
  const int nb = num_base_params(base);
  if (nb == 0 || tc < 0 || tc > 2) return 0;
  return nb + (tc == TC_NONE ? 0 : 4);
}

// Per-parameter-set constants (everything that depends on theta only).
struct ModelConsts {
  double sigma, hs2;          // sigma, sigma^2/2
  double comp;                // jump compensator per unit time: lambda(E e^J - 1) or k(1)
  double lam, muJ, hsJ2;      // Merton
  double cg, G, M, Y, myGy;   // CGMY: cg = C Gamma(-Y), myGy = M^Y + G^Y;  VG: cg = 1/nu, G = theta nu, M = sigma^2 nu/2
  double speed, adaV, v0, aos2, kap1;  // CIR: aos2 = speed/adaV^2, kappa(u) = speed - kap1 u
};

__host__ __device__ __forceinline__ void make_consts(int base, int tc, const double* __restrict__ p,
                                            ModelConsts& c) {
  c.sigma = p[0];
  c.hs2 = 0.5 * p[0] * p[0];
  c.comp = 0.0;
  c.lam = c.muJ = c.hsJ2 = 0.0;
  c.cg = c.G = c.M = c.Y = c.myGy = 0.0;
  if (base == BASE_MERTON) {
    c.lam = p[1];
    c.muJ = p[2];
    c.hsJ2 = 0.5 * p[3] * p[3];
    c.comp = c.lam * (exp(c.muJ + c.hsJ2) - 1.0);
  } else if (base == BASE_CGMY) {
    const double C = p[1];
    c.G = p[2];
    c.M = p[3];
    c.Y = p[4];
    if (C != 0.0) {
      c.cg = C * tgamma(-c.Y);
      c.myGy = pow(c.M, c.Y) + pow(c.G, c.Y);
      c.comp = c.cg * (pow(c.M - 1.0, c.Y) + pow(c.G + 1.0, c.Y) - c.myGy);
    }
  } else if (base == BASE_VG) {  // params [sigma, theta, nu]; pure jump: no Brownian part
    c.cg = 1.0 / p[2];
    c.G = p[1] * p[2];
    c.M = 0.5 * p[0] * p[0] * p[2];
    c.comp = -c.cg * log(1.0 - c.G - c.M);
    c.sigma = 0.0;
    c.hs2 = 0.0;
  }
  c.speed = c.adaV = c.v0 = c.aos2 = c.kap1 = 0.0;
  if (tc != TC_NONE) {
    const double* q = p + num_base_params(base);
    c.speed = q[0];
    c.adaV = q[1];
    c.v0 = q[3];
    c.kap1 = q[1] * q[2] * c.sigma;  // adaV * rho * sigma
    c.aos2 = c.speed / (c.adaV * c.adaV);
  }
}

// uncompensated jump exponent
__host__ __device__ __forceinline__ cd jump_exponent(int base, const ModelConsts& c, cd u) {
  if (base == BASE_MERTON) {
    // lambda (exp(u muJ + sJ^2 u^2/2) - 1)
    const cd uu = u * u;
    const cd e = cexp(mk(u.x * c.muJ + c.hsJ2 * uu.x, u.y * c.muJ + c.hsJ2 * uu.y));
    return mk(c.lam * (e.x - 1.0), c.lam * e.y);
  }
  if (base == BASE_CGMY) {
    if (c.cg == 0.0) return mk(0.0, 0.0);
    const cd a = cpow(mk(c.M - u.x, -u.y), c.Y);
    const cd b = cpow(mk(c.G + u.x, u.y), c.Y);
    return mk(c.cg * (a.x + b.x - c.myGy), c.cg * (a.y + b.y));
  }
  if (base == BASE_VG) {  // -(1/nu) log(1 - u theta nu - sigma^2 nu u^2 / 2)
    const cd uu = u * u;
    const cd l = clog(mk(1.0 - u.x * c.G - c.M * uu.x, -u.y * c.G - c.M * uu.y));
    return mk(-c.cg * l.x, -c.cg * l.y);
  }
  return mk(0.0, 0.0);
}

// risk-neutral Levy exponent per unit time: u (rate - s2 - comp) + s2 u^2 + jumps(u),
// s2 = sigma^2/2 (or 0 when the diffusion is switched off)
__host__ __device__ __forceinline__ cd levy_exponent(int base, const ModelConsts& c, cd u, double rate,
                                            double s2, bool with_jumps) {
  const double mu = rate - s2 - (with_jumps ? c.comp : 0.0);
  const cd uu = u * u;
  cd r = mk(u.x * mu + s2 * uu.x, u.y * mu + s2 * uu.y);
  if (with_jumps && base != BASE_GAUSS) r = r + jump_exponent(base, c, u);
  return r;
}

// T-independent part of log phi(u)
struct CfPre {
  cd lin;    // log phi += lin * T
  cd pod;    // psi/delta            (adaV == 0: psi)
  cd dmk;    // delta - kappa
  cd g;      // (delta-kappa)/(2 delta)
  cd delta;  //                       (adaV == 0: kappa)
};

__host__ __device__ __forceinline__ void cf_prepare(int base, int tc, const ModelConsts& c, cd u, double r,
                                           CfPre& pre) {
  if (tc == TC_NONE) {
    pre.lin = levy_exponent(base, c, u, r, c.hs2, true);
    return;
  }
  cd psi;  // argument of cirLogMGF = -(time-changed exponent with r = 0)
  if (tc == TC_CIR) {
    psi = -levy_exponent(base, c, u, 0.0, c.hs2, true);
    pre.lin = mk(r * u.x, r * u.y);
  } else {  // only the diffusion is time changed; jumps (compensated, sigma = 0) in calendar time
    psi = -levy_exponent(BASE_GAUSS, c, u, 0.0, c.hs2, false);
    pre.lin = mk(r * u.x, r * u.y) + levy_exponent(base, c, u, 0.0, 0.0, true);
  }
  const cd kap = mk(c.speed - c.kap1 * u.x, -c.kap1 * u.y);
  if (c.adaV == 0.0) {
    pre.pod = psi;
    pre.delta = kap;
    return;
  }
  const double s2v = 2.0 * c.adaV * c.adaV;
  const cd kk = kap * kap;
  const cd delta = csqrt_fast(mk(kk.x + s2v * psi.x, kk.y + s2v * psi.y));
  pre.delta = delta;
  pre.dmk = delta - kap;
  // g = (delta - kappa)/(2 delta) and psi/delta share the reciprocal of |delta|^2
  const double invd = fm_rcp(fma(delta.x, delta.x, delta.y * delta.y));
  const cd cdl = mk(delta.x * invd, -delta.y * invd);          // 1 / delta
  const cd gd = pre.dmk * cdl;
  pre.g = mk(0.5 * gd.x, 0.5 * gd.y);
  pre.pod = psi * cdl;
}

// log phi(u) at maturity T, straight-line variants (no data-dependent branch) so that several
// maturities of one (set, u) pair interleave:
//   CF_MODE_LEVY  no time change          CF_MODE_CIR  CIR clock, adaV != 0
//   CF_MODE_DET   deterministic clock (adaV == 0, test.cpp:1187)
enum { CF_MODE_LEVY = 0, CF_MODE_CIR = 1, CF_MODE_DET = 2 };
__host__ __device__ __forceinline__ int cf_mode(int tc, const ModelConsts& c) {
  return tc == TC_NONE ? CF_MODE_LEVY : (c.adaV == 0.0 ? CF_MODE_DET : CF_MODE_CIR);
}
// CIR clock, given e = exp(-delta T) (computed directly, or as a power of exp(-delta dt) when the
// maturities sit on a common grid, see coeff_maturities) and l = lin * T
template <class T>
__host__ __device__ __forceinline__ cx<T> cf_log_cir_given_e(const ModelConsts& c, const CfPre& pre, T Tm,
                                                    cx<T> l, cx<T> e) {
  const cx<T> ome = 1.0 - e;
  const cx<T> w = 1.0 - pre.g * ome;
  cx<T> q, lw;
  cdiv_clog_unit(ome, w, &q, &lw);
  const cx<T> b = pre.pod * q;
  const cx<T> cc = mk(c.aos2 * (2.0 * lw.x + pre.dmk.x * Tm), c.aos2 * (2.0 * lw.y + pre.dmk.y * Tm));
  return l - (c.v0 * b) - cc;
}

// T = double, or VD<N>: N maturities of the same (set, u) pair advanced operation by operation
template <int MODE, class T>
__host__ __device__ __forceinline__ cx<T> cf_log_at_mode(const ModelConsts& c, const CfPre& pre, T Tm) {
  const cx<T> l = mk(pre.lin.x * Tm, pre.lin.y * Tm);
  if (MODE == CF_MODE_LEVY) return l;
  if (MODE == CF_MODE_DET) {
    const cd kp = pre.delta, ps = pre.pod;
    const cx<T> ek = 1.0 - cexp(mk(-kp.x * Tm, -kp.y * Tm));
    const cx<T> ekk = cdiv(ek, kp);
    const cx<T> b = ps * ekk;
    const cx<T> cc = cdiv(mk(c.speed * ps.x, c.speed * ps.y), kp) * (Tm - ekk);
    return l - (c.v0 * b) - cc;
  }
  return cf_log_cir_given_e(c, pre, Tm, l, cexp(mk(-pre.delta.x * Tm, -pre.delta.y * Tm)));
}
__host__ __device__ __forceinline__ cd cf_log_at(int tc, const ModelConsts& c, const CfPre& pre, double T) {
  const int mode = cf_mode(tc, c);
  if (mode == CF_MODE_LEVY) return cf_log_at_mode<CF_MODE_LEVY>(c, pre, T);
  if (mode == CF_MODE_DET) return cf_log_at_mode<CF_MODE_DET>(c, pre, T);
  return cf_log_at_mode<CF_MODE_CIR>(c, pre, T);
}

// ---- re-associated CIR clock for the coefficient kernels ---------------------------------------
// log phi = lin T - v0 b - aos2 (2 log w + (delta - kappa) T),  b = (psi/delta) (1 - e)/w,
// w = 1 - g (1 - e), e = exp(-delta T)   [the same formula as cf_log_cir_given_e], with everything
// that does not depend on T folded once per (set, u) pair:
//   linA = lin - aos2 (delta - kappa)        pv = v0 psi/delta        gm1 = 1 - g
//   log phi = linA T - pv (1 - e)/w - aos2 (log|w|^2, 2 arg w)
// 124 instead of 150 fp64 instructions per maturity; the fp64 pipe issues at half rate, so every
// fp64 instruction saved is worth two integer / select instructions (profiles/r2_coeff_*).
struct CirFast {
  cd linA, pv, g, gm1;
  double aos2, aos2x2;
};
__host__ __device__ __forceinline__ CirFast cir_fast_prepare(const ModelConsts& c, const CfPre& pre) {
  CirFast f;
  f.aos2 = c.aos2;
  f.aos2x2 = 2.0 * c.aos2;
  f.linA = mk(fma(-c.aos2, pre.dmk.x, pre.lin.x), fma(-c.aos2, pre.dmk.y, pre.lin.y));
  f.pv = mk(c.v0 * pre.pod.x, c.v0 * pre.pod.y);
  f.g = pre.g;
  f.gm1 = mk(1.0 - pre.g.x, -pre.g.y);
  return f;
}
// log phi at maturity Tm given e = exp(-delta Tm)
template <class T>
__host__ __device__ __forceinline__ cx<T> cir_logphi_fast(const CirFast& f, T Tm, cx<T> e) {
  // w = (1 - g) + g e
  const cx<T> w = mk(fm_fma(f.g.x, e.x, fm_fma(-f.g.y, e.y, f.gm1.x)),
                     fm_fma(f.g.x, e.y, fm_fma(f.g.y, e.x, f.gm1.y)));
  const T n = fm_fma(w.x, w.x, w.y * w.y);
  const T inv = fm_rcp(n);
  // q = (1 - e) conj(w) / |w|^2
  const T omx = 1.0 - e.x;                       // ome = (omx, -e.y)
  const T qx = fm_fma(omx, w.x, -(e.y * w.y)) * inv;
  const T qy = -fm_fma(e.y, w.x, omx * w.y) * inv;
  const T L = fm_log_pos(n);                     // log |w|^2
  const T th = fm_atan2_normal(w.y, w.x);        // arg w
  // pv q + aos2 (L, 2 th)
  const T tx = fm_fma(f.pv.x, qx, fm_fma(-f.pv.y, qy, f.aos2 * L));
  const T ty = fm_fma(f.pv.x, qy, fm_fma(f.pv.y, qx, f.aos2x2 * th));
  return mk(fm_fma(f.linA.x, Tm, -tx), fm_fma(f.linA.y, Tm, -ty));
}

// Option transforms applied to phi(u) (OptionPricing.h:44-73): 0 price, 1 delta, 2 gamma, 3 theta
template <class T>
__host__ __device__ __forceinline__ cx<T> option_transform(int greek, cx<T> phi, cd u, double r) {
  if (greek == 0) return phi;
  if (greek == 1) return phi * u;
  if (greek == 2) return -(phi * (u * (1.0 - u)));
  // theta; log of the CF VALUE (principal branch), as the reference does; select, not branch
  const cx<T> l = clog(phi);
  const cx<T> t = -(mk(l.x - r, l.y) * phi);
  const auto pos = phi.x > 0.0;
  return mk(fm_sel(pos, t.x, 0.0), fm_sel(pos, t.y, 0.0));
}

}  // namespace fop
