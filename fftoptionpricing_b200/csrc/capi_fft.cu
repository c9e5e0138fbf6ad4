// C ABI: fop_fft, fop_carr_madan  (include/fftop_b200.h)
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cf_models.cuh"
#include "fft_engine.cuh"
#include "fft_run.cuh"
#include "handle.cuh"

namespace fop {
// cached exact twiddle table e^{-2 pi i k/N} (also used by capi_fsts.cu)
int get_twiddles_shared(fop_handle_t h, int N, const double2** out) {
  for (auto& t : h->twiddles)
    if (t.N == N) {
      *out = t.d;
      return FOP_OK;
    }
  double2* d = nullptr;
  FOP_CUDA(h, cudaMalloc(&d, sizeof(double2) * (size_t)N));
  twiddle_table_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(d, N);
  FOP_LAUNCH_CHECK(h);
  h->twiddles.push_back({N, d});
  *out = d;
  return FOP_OK;
}
}  // namespace fop
namespace {
// direct O(N^2) transform for N < 16 (one thread per output)
__global__ void tiny_dft_kernel(cd* data, int batch, int N, int sign) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch) return;
  cd x[8], y[8];
  for (int j = 0; j < N; ++j) x[j] = data[(size_t)idx * N + j];
  for (int k = 0; k < N; ++k) {
    cd acc = mk(0.0, 0.0);
    for (int j = 0; j < N; ++j) {
      double s, c;
      sincospi(sign * 2.0 * (double)((j * k) % N) / (double)N, &s, &c);
      acc = acc + x[j] * mk(c, s);
    }
    y[k] = acc;
  }
  for (int k = 0; k < N; ++k) data[(size_t)idx * N + k] = y[k];
}

// ---- Carr-Madan functors (OptionPricing.h:564-567 CallAug, :582-607 CarrMadanGeneric)
// Exact-zero shortcut.  For a Levy model with a diffusion part, |phi(a + i w)| <= phi(a) e^{-s2 T w^2}
// (s2 = sigma^2/2; the jump exponent satisfies Re J(a + i w) <= J(a) because |e^{(a + i w) x}| =
// e^{a x}).  Where this bound says Re log phi < -750 the reference's own exp() underflows to
// exactly 0 (below e^{-745.14}) and so does x_j: the point is not evaluated.  On config 4
// (N = 65536, ada = 0.25: w up to 16384) that is ~95 % of the grid.
struct CmAux {
  ModelConsts mc;
  double c0, s2t;  // bound: Re log phi(a + i w) <= c0 - s2t w^2   (s2t = 0: never skip)
};
struct CmGen {
  static constexpr bool has_cf = true;
  // per-set constants of the transforms a CTA works on at a time
  static size_t aux_bytes(int sets_per_cta) { return sizeof(CmAux) * (size_t)sets_per_cta; }
  int base, tc, np;
  const double* params;
  double r, T, ada, alpha, discount, b;
  int no_skip;      // FOP_CM_NO_SKIP=1: evaluate every point (A/B runs, tests)
  const int* idx;   // optional: transform p works on parameter set idx[p]
  // optional: per-set constants already formed by cm_cutoff_kernel (one thread per set, all sets in
  // parallel) -- otherwise ONE thread of each CTA forms them (tgamma, four pows for CGMY) while the
  // other 255 wait at the barrier: a third of cm_head_kernel's warp samples (profiles/r2_all_kernels)
  const CmAux* aux_all = nullptr;
  __device__ void fill(CmAux& me, int p) const {
    if (aux_all) {
      me = aux_all[idx ? idx[p] : p];
      return;
    }
    make_consts(base, tc, params + (size_t)(idx ? idx[p] : p) * np, me.mc);
    me.c0 = 0.0;
    me.s2t = 0.0;
    if (tc == TC_NONE && me.mc.hs2 > 0.0 && !no_skip) {
      const double a = alpha + 1.0;
      const cd ja = jump_exponent(base, me.mc, mk(a, 0.0));  // real when a lies in the CF's strip
      const double c0 = T * (a * (r - me.mc.hs2 - me.mc.comp) + me.mc.hs2 * a * a + ja.x);
      if (ja.y == 0.0 && fabs(c0) < 1e300) {  // false for NaN
        me.c0 = c0;
        me.s2t = me.mc.hs2 * T;
      }
    }
  }
  // the predicate of the shortcut, shared with the cut-off kernel of the pruned path
  __device__ static bool is_zero(const CmAux& ax, double w) { return fma(-ax.s2t, w * w, ax.c0) < -750.0; }
  __device__ void prepare(void* aux, int p0, int count) const {
    CmAux* ax = reinterpret_cast<CmAux*>(aux);
    if ((int)threadIdx.x < count) fill(ax[threadIdx.x], p0 + threadIdx.x);
    __syncthreads();
  }
  // x_j = discount * phi(v + alpha + 1) / (alpha^2 + alpha + v^2 + (2 alpha + 1) v) * e^{i b j ada}
  //       * (3 -+ 1),  v = i j ada;  x_0 halved            (OptionPricing.h:594-599)
  __device__ cd operator()(const void* aux, int p0, int p, int j) const {
    const CmAux& ax = reinterpret_cast<const CmAux*>(aux)[p - p0];
    const ModelConsts& mc = ax.mc;
    const double w = j * ada;
    if (is_zero(ax, w)) return mk(0.0, 0.0);
    const cd u = mk(alpha + 1.0, w);
    CfPre pre;
    cf_prepare(base, tc, mc, u, r, pre);
    const cd phi = cexp(cf_log_at(tc, mc, pre, T));
    const cd den = mk(alpha * alpha + alpha - w * w, (2.0 * alpha + 1.0) * w);
    const cd aug = cdiv(phi, den);
    const cd ph = cis(b * j * ada);  // evaluated on the rounded product, like the reference
    const double wt = j == 0 ? 1.0 : ((j & 1) ? 4.0 : 2.0);
    const cd v = aug * ph;
    const double sc = discount * wt;
    return mk(v.x * sc, v.y * sc);
  }
};
struct CmEpi {
  double* out;
  int N;
  double S0, alpha, b, lambda, ada, discount;
  int is_put;
  const int* idx;      // optional: transform p writes the row of parameter set idx[p]
  // (per-index tables of the two exps, shared by the batch, were tried: the strided look-ups of
  // the pruned path cost more L2 sectors than the exps cost fp64 issue slots -- 15.3 vs 11.3 ms)
  __device__ double undamp(int k) const { return fm_exp(-alpha * (-b + lambda * k)); }
  __device__ double strike_over_s0(int k) const { return fm_exp(-b + lambda * k); }
  // call_m = S0 Re(X_m) e^{-alpha(-b + lambda m)} ada/(3 pi)  (:605);  put by parity (:652)
  __device__ double price(int k, double re) const {
    double v = S0 * re * undamp(k) * ada / (M_PI * 3.0);
    if (is_put) v = v - S0 + S0 * strike_over_s0(k) * discount;
    return v;
  }
  __device__ void operator()(int p, int k, cd X) const {
    out[(size_t)(idx ? idx[p] : p) * N + k] = price(k, X.x);
  }
};
// ---- input-pruned Carr-Madan for long grids ----------------------------------------------------
// When the shortcut above proves x_j = 0 for every j >= L = N / R, the N-point transform is R
// independent L-point transforms of the pre-twiddled head, one per residue class of the output:
//   X_{R q + s} = sum_{j < L} (x_j W_N^{j s}) W_L^{j q},   s = 0 .. R-1.
// L <= 8192 fits the single-CTA shared-memory FFT, so the four-step transposition (16 B per point
// written and re-read) disappears; the head x_j is generated once (cm_head_kernel) and re-read by
// the R transforms of its set from L2.
__global__ void cm_cutoff_kernel(CmGen g, int P, int N, int* jc_out, CmAux* aux_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  CmAux ax;
  g.fill(ax, p);
  aux_out[p] = ax;
  int jc = N;
  if (ax.s2t > 0.0) {
    const double w2 = fmax((ax.c0 + 750.0) / ax.s2t, 0.0);
    const double je = sqrt(w2) / g.ada;
    jc = je < (double)N ? (int)je : N;
    // exact consistency with the generator's own predicate (monotone in j)
    while (jc > 0 && CmGen::is_zero(ax, (jc - 1) * g.ada)) --jc;
    while (jc < N && !CmGen::is_zero(ax, jc * g.ada)) ++jc;
  }
  jc_out[p] = jc;  // x_j = 0 for all j >= jc
}
// x[t][j] for the first L grid points of transform t: one CTA per set (the per-set constants --
// tgamma / pow for CGMY -- are formed once), threads stride over the head
__global__ void __launch_bounds__(256) cm_head_kernel(CmGen g, int L, cd* __restrict__ xbuf) {
  __shared__ CmAux ax;
  const int t = blockIdx.x;
  if (threadIdx.x == 0) g.fill(ax, t);
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += 256) xbuf[(size_t)t * L + j] = g(&ax, t, t, j);
}
// Two residue classes per complex transform: only Re X is needed (OptionPricing.h:605), and for
// h_j = (y_j + conj(y_{(L-j) mod L}))/2 the transform of h is exactly Re Y.  With y_j = x_j W_N^{js},
// conj(y_{L-j}) = conj(x_{L-j}) conj(W_R^s) W_N^{js}, so
//   z_j = h^{(s)}_j + i h^{(s+1)}_j
//       = W_N^{js}/2 [x_j + c_s conj(x_{L-j})] + i W_N^{j(s+1)}/2 [x_j + c_{s+1} conj(x_{L-j})],
//   c_s = e^{+2 pi i s/R};   Z_q = Re Y^{(s)}_q + i Re Y^{(s+1)}_q  ->  outputs R q + s, R q + s + 1
// (adjacent: one 16-byte store).  Transform p of the batch = (set p >> (logR-1), pair p & (R/2-1)).
struct PrunedGen {
  static constexpr bool has_cf = false;  // loads and a few multiplications
  static size_t aux_bytes(int) { return 0; }
  const cd* xbuf;
  const double2* twN;
  int L, logR, N;
  __device__ void prepare(void*, int, int) const {}
  __device__ cd operator()(const void*, int, int p, int j) const {
    const int t = p >> (logR - 1), s = 2 * (p & ((1 << (logR - 1)) - 1));
    const cd* xs = xbuf + (size_t)t * L;
    const cd x = xs[j];
    if (j == 0) return mk(x.x, x.x);  // the self-conjugate point: Re y_0 for both residues
    const cd xr = conj(xs[L - j]);
    const cd w1 = tw_load<-1>(twN, (j * s) & (N - 1));
    const cd w2 = w1 * tw_load<-1>(twN, j);
    const cd a = w1 * (x + tw_load<1>(twN, s * L) * xr);
    const cd b = w2 * (x + tw_load<1>(twN, (s + 1) * L) * xr);
    return mk(0.5 * (a.x - b.y), 0.5 * (a.y + b.x));
  }
};
struct PrunedEpi {
  CmEpi inner;
  int logR, vec2;
  double ustep, kstep;  // e^{-alpha lambda 256 R}, e^{lambda 256 R}: factor steps of the 4096-point kernel's epilogue
  __device__ void operator()(int p, int k, cd Z) const {
    const int t = p >> (logR - 1), s = 2 * (p & ((1 << (logR - 1)) - 1));
    const int m = (k << logR) + s;
    const double v0 = inner.price(m, Z.x), v1 = inner.price(m + 1, Z.y);
    double* o = inner.out + (size_t)(inner.idx ? inner.idx[t] : t) * inner.N + m;
    if (vec2) {
      *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
    } else {
      o[0] = v0;
      o[1] = v1;
    }
  }
};

// Dedicated L = 4096 = 16^3 version of the pruned transform (87 % of the sets of config 4): the
// generator feeds the first radix-16 pass in registers and the last pass feeds the epilogue from
// registers, so the grid makes two shared-memory round trips instead of four and the digit-reversed
// (bank-conflicting) epilogue read disappears.  One transform (= two residue classes of one set)
// per CTA of 256 threads.
__global__ void __launch_bounds__(256, 2) cm_pruned4096_kernel(PrunedGen pg, PrunedEpi pe, int batch,
                                                               const double2* __restrict__ twL) {
  constexpr int n = 12, L = 1 << n, S0 = L >> 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* sm = reinterpret_cast<cd*>(smem_raw);
  const int b = threadIdx.x;
  // thread-independent twiddles of the generator in shared memory for the whole launch (they were 35
  // dependent table reads per thread and transform; ncu: a third of the warp samples waited on the
  // generator's global loads): twA[pair][t] = W_N^{256 t (2 pair)}, twB[t] = W_N^{256 t}, twC[s] = c_s
  const int npairs = 1 << (pg.logR - 1);
  cd* twA = sm + fft_padded_size(L);
  cd* twB = twA + (size_t)npairs * 16;
  cd* twC = twB + 16;
  for (int i = b; i < npairs * 16; i += 256) twA[i] = tw_load<-1>(pg.twN, (256 * (i & 15) * 2 * (i >> 4)) & (pg.N - 1));
  if (b < 16) twB[b] = tw_load<-1>(pg.twN, 256 * b);
  for (int i = b; i < 2 * npairs; i += 256) twC[i] = tw_load<1>(pg.twN, i * pg.L);
  __syncthreads();
  const cd w0 = tw_load<-1>(twL, b);               // W_L^{b}: pass 0 (stride 256), q = b
  const cd w1 = tw_load<-1>(twL, (b & 15) * 16);   // W_256^{b mod 16}: pass 1 (stride 16)
  const int base1 = fft_base16(b, 16);
  // twiddles of the generator without per-point table traffic: j = b + 256 t, so
  //   W_N^{j s} = W_N^{b s} W_256^{t s'},  W_N^{j} = W_N^{b} W_256^{t'}   (W_N^{256 m} = W_{N/256}^m)
  // one strided table read per transform for W_N^{b s}; the 256 / (N/256)-periodic factors come
  // from a table slice that stays in L1
  const int N = pg.N, L_ = pg.L, logR = pg.logR;
  const cd wb = tw_load<-1>(pg.twN, b);
  // (consecutive CTAs take the R/2 transforms of one set: their 16-byte stores interleave into full
  // sectors in L2 while they are in flight together -- one CTA doing them in turn measured 12.7 ms
  // against 8.0 ms)
  for (int p = blockIdx.x; p < batch; p += gridDim.x) {
    const int tset = p >> (logR - 1), s = 2 * (p & ((1 << (logR - 1)) - 1));
    const cd* xs = pg.xbuf + (size_t)tset * L_;
    const cd wbs = tw_load<-1>(pg.twN, (b * s) & (N - 1));
    const cd cs0 = twC[s], cs1 = twC[s + 1];
    const cd* twAp = twA + (size_t)(s >> 1) * 16;
    cd x[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int j = b + S0 * t;
      const cd xv = xs[j];
      if (j == 0) {
        x[t] = mk(xv.x, xv.x);
        continue;
      }
      const cd xr = conj(xs[L_ - j]);
      const cd w1 = wbs * twAp[t];
      const cd w2 = w1 * (wb * twB[t]);
      const cd a = w1 * (xv + cs0 * xr);
      const cd bb = w2 * (xv + cs1 * xr);
      x[t] = mk(0.5 * (a.x - bb.y), 0.5 * (a.y + bb.x));
    }
    dif16_regs_pow<-1>(x, w0);
    store16_dft<-1>(sm, b, S0, x);
    __syncthreads();
    load16(sm, base1, 16, x);
    dif16_regs_pow<-1>(x, w1);
    store16_dft<-1>(sm, base1, 16, x);
    __syncthreads();
    load16(sm, 16 * b, 1, x);
    Dft<16, -1>::run(x);
    // position 16 b + t holds frequency k = (b >> 4) + 16 (b & 15) + 256 t  (digit reversal), i.e.
    // outputs m = R (k0 + 256 t) + s, + 1.  The undamping factor e^{-alpha(-b + lambda m)} of
    // OptionPricing.h:605 (and the strike e^{-b + lambda m} of the put, :652) factor over (k0, t, s):
    // one exp per thread for k0 (outside the transform loop), a running product over t, two exps
    // per transform for s, s + 1 -- instead of one or two exps per output (which were ~40 % of the
    // kernel's fp64 instructions).  The running product carries <= 16 roundings (1e-15 relative).
    const int k0 = (b >> 4) + 16 * (b & 15);
    const CmEpi& ce = pe.inner;
    const double al = ce.alpha * ce.lambda;
    const double f0 = fm_exp(-al * s), f1 = fm_exp(-al * (s + 1));
    const double rk0 = (double)((long long)k0 << logR);
    double ut = ce.S0 * fm_exp(-ce.alpha * (-ce.b + ce.lambda * rk0)) * ce.ada / (M_PI * 3.0);
    double kt = 0.0, g0 = 0.0, g1 = 0.0;
    if (ce.is_put) {
      g0 = fm_exp(ce.lambda * s);
      g1 = fm_exp(ce.lambda * (s + 1));
      kt = ce.S0 * fm_exp(-ce.b + ce.lambda * rk0) * ce.discount;
    }
    const double ustep = pe.ustep, kstep = pe.kstep;
    double* orow = ce.out + (size_t)(ce.idx ? ce.idx[tset] : tset) * ce.N + ((size_t)k0 << logR) + s;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const cd Z = x[Dft<16, -1>::out(t)];
      double v0 = Z.x * (ut * f0), v1 = Z.y * (ut * f1);
      if (ce.is_put) {
        v0 = v0 - ce.S0 + kt * g0;
        v1 = v1 - ce.S0 + kt * g1;
      }
      double* o = orow + ((size_t)(256 * t) << logR);
      if (pe.vec2) {
        *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
      } else {
        o[0] = v0;
        o[1] = v1;
      }
      ut *= ustep;
      kt *= kstep;
    }
    __syncthreads();  // the next transform's first pass overwrites the grid
  }
}

// Set-per-CTA version of the L = 4096 pruned transform (FOP_CM_PRUNED=set; an experiment that did not
// beat the kernel above, kept selectable with its evidence).  ncu on cm_pruned4096_kernel (profiles/r2_cm_pruned_ncu.json): 54 % of
// the warp samples wait on the global loads of the generator -- every one of the R/2 transforms of
// a set re-reads the whole 64 KB head (x_j and x_{L-j}) from L2 behind dependent address
// arithmetic -- and the head itself makes an HBM round trip out of cm_head_kernel.  Here one
// persistent CTA of 512 threads owns a set: it evaluates the head ONCE into shared memory (the
// cm_head_kernel arithmetic, no HBM buffer), then its two groups of 256 threads run the R/2
// transforms two at a time out of shared memory (own FFT grid each, named barriers per group).  In
// a round the groups produce residues 4r .. 4r + 3 of every frequency: the two 16-byte stores
// complete a 32-byte sector together.
// out of line: the CF evaluation gets its own register allocation instead of competing with the
// transform's 16-point register tile
__device__ __noinline__ void cm_head_fill(const CmGen& g, const CmAux* ax, int set, cd* head, int L) {
  for (int j = threadIdx.x; j < L; j += blockDim.x) head[j] = g(ax, set, set, j);
}
__global__ void __launch_bounds__(512, 1) cm_pruned4096_set_kernel(CmGen g, PrunedEpi pe, int nsets, int N,
                                                                   const double2* __restrict__ twN,
                                                                   const double2* __restrict__ twL) {
  constexpr int n = 12, L = 1 << n, S0 = L >> 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* head = reinterpret_cast<cd*>(smem_raw);
  const int grp = threadIdx.x >> 8, b = threadIdx.x & 255;
  cd* sm = head + L + (size_t)grp * fft_padded_size(L);
  const int logR = pe.logR, npairs = 1 << (logR - 1);
  // The generator's twiddle factors that do not depend on the thread -- W_N^{256 t s}, W_N^{256 t},
  // c_s, c_{s+1} -- sit in shared memory for the whole launch: with ~200 KB of shared memory in use the
  // L1 is a few KB, and the per-point table reads were L2 round trips (47 % long-scoreboard stalls).
  cd* twA = head + L + 2 * (size_t)fft_padded_size(L);  // [npairs][16]: W_N^{256 t (2 pair)}
  cd* twB = twA + (size_t)npairs * 16;                  // [16]:         W_N^{256 t}
  cd* twC = twB + 16;                                   // [npairs][2]:  c_s, c_{s+1}
  CmAux* ax = reinterpret_cast<CmAux*>(twC + 2 * (size_t)npairs);
  for (int i = threadIdx.x; i < npairs * 16; i += 512)
    twA[i] = tw_load<-1>(twN, (256 * (i & 15) * 2 * (i >> 4)) & (N - 1));
  if (threadIdx.x < 16) twB[threadIdx.x] = tw_load<-1>(twN, 256 * threadIdx.x);
  for (int i = threadIdx.x; i < 2 * npairs; i += 512) twC[i] = tw_load<1>(twN, i * L);
  const cd w0 = tw_load<-1>(twL, b);
  const cd w1 = tw_load<-1>(twL, (b & 15) * 16);
  const int base1 = fft_base16(b, 16);
  const cd wb = tw_load<-1>(twN, b);
  const int k0 = (b >> 4) + 16 * (b & 15);
  const CmEpi& ce = pe.inner;
  auto group_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory"); };
  for (int set = blockIdx.x; set < nsets; set += gridDim.x) {
    __syncthreads();  // the previous set's transforms have finished reading the head
    if (threadIdx.x == 0) g.fill(*ax, set);
    __syncthreads();
    cm_head_fill(g, ax, set, head, L);
    __syncthreads();
    double* orow_set = ce.out + (size_t)(ce.idx ? ce.idx[set] : set) * ce.N;
    const double rk0 = (double)((long long)k0 << logR);
    for (int pair = grp; pair < npairs; pair += 2) {
      const int s = 2 * pair;
      const cd wbs = tw_load<-1>(twN, (b * s) & (N - 1));
      const cd cs0 = twC[s], cs1 = twC[s + 1];
      cd x[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int j = b + S0 * t;
        const cd xv = head[j];
        if (j == 0) {
          x[t] = mk(xv.x, xv.x);
          continue;
        }
        const cd xr = conj(head[L - j]);
        const cd w1t = wbs * twA[pair * 16 + t];
        const cd w2t = w1t * (wb * twB[t]);
        const cd a = w1t * (xv + cs0 * xr);
        const cd bb = w2t * (xv + cs1 * xr);
        x[t] = mk(0.5 * (a.x - bb.y), 0.5 * (a.y + bb.x));
      }
      dif16_regs_pow<-1>(x, w0);
      store16_dft<-1>(sm, b, S0, x);
      group_sync();
      load16(sm, base1, 16, x);
      dif16_regs_pow<-1>(x, w1);
      store16_dft<-1>(sm, base1, 16, x);
      group_sync();
      load16(sm, 16 * b, 1, x);
      Dft<16, -1>::run(x);
      // epilogue: factored undamping, see cm_pruned4096_kernel
      const double al = ce.alpha * ce.lambda;
      const double f0 = fm_exp(-al * s), f1 = fm_exp(-al * (s + 1));
      double ut = ce.S0 * fm_exp(-ce.alpha * (-ce.b + ce.lambda * rk0)) * ce.ada / (M_PI * 3.0);
      double kt = 0.0, g0 = 0.0, g1 = 0.0;
      if (ce.is_put) {
        g0 = fm_exp(ce.lambda * s);
        g1 = fm_exp(ce.lambda * (s + 1));
        kt = ce.S0 * fm_exp(-ce.b + ce.lambda * rk0) * ce.discount;
      }
      double* orow = orow_set + ((size_t)k0 << logR) + s;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const cd Z = x[Dft<16, -1>::out(t)];
        double v0 = Z.x * (ut * f0), v1 = Z.y * (ut * f1);
        if (ce.is_put) {
          v0 = v0 - ce.S0 + kt * g0;
          v1 = v1 - ce.S0 + kt * g1;
        }
        double* o = orow + ((size_t)(256 * t) << logR);
        if (pe.vec2) {
          *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
        } else {
          o[0] = v0;
          o[1] = v1;
        }
        ut *= pe.ustep;
        kt *= pe.kstep;
      }
      group_sync();  // the group's next transform overwrites its grid
    }
  }
}

// Warp-specialised version (FOP_CM_PRUNED=ws; experiment): the same set-per-CTA scheme, but the two halves of the CTA
// take different ROLES -- 8 producer warps evaluate the head of set i + 1 into the second of two
// head buffers while 8 consumer warps run the R/2 transforms of set i.  The two phases want
// different resources (CF evaluation: the fp64 dispatch port; transforms: shared memory, barriers
// and the scattered 16-byte stores), so side by side they fill each other's stalls instead of
// alternating as in cm_pruned4096_set_kernel (ncu: 44 % long-scoreboard + 13 % barrier stalls with
// the phases serialised).  Hand-off through named barriers (full / empty per buffer).
__global__ void __launch_bounds__(512, 1) cm_pruned4096_ws_kernel(CmGen g, PrunedEpi pe, int nsets, int N,
                                                                  const double2* __restrict__ twN,
                                                                  const double2* __restrict__ twL) {
  constexpr int n = 12, L = 1 << n, S0 = L >> 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* head = reinterpret_cast<cd*>(smem_raw);                    // [2][L]
  cd* sm = head + 2 * L;                                         // FFT grid (padded)
  const int logR = pe.logR, npairs = 1 << (logR - 1);
  cd* twA = sm + fft_padded_size(L);                             // [npairs][16]: W_N^{256 t (2 pair)}
  cd* twB = twA + (size_t)npairs * 16;                           // [16]:         W_N^{256 t}
  cd* twC = twB + 16;                                            // [npairs][2]:  c_s, c_{s+1}
  CmAux* ax = reinterpret_cast<CmAux*>(twC + 2 * (size_t)npairs);  // [2]
  const int role = threadIdx.x >> 8, b = threadIdx.x & 255;
  for (int i = threadIdx.x; i < npairs * 16; i += 512)
    twA[i] = tw_load<-1>(twN, (256 * (i & 15) * 2 * (i >> 4)) & (N - 1));
  if (threadIdx.x < 16) twB[threadIdx.x] = tw_load<-1>(twN, 256 * threadIdx.x);
  for (int i = threadIdx.x; i < 2 * npairs; i += 512) twC[i] = tw_load<1>(twN, i * L);
  __syncthreads();
  const int nloc = (nsets - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // named barriers: 1, 2 = full[buf] (256 arrive + 256 sync), 3, 4 = empty[buf], 5 = producers, 6 = consumers
  auto bar_sync = [](int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); };
  auto bar_arrive = [](int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); };
  if (role == 1) {  // ---- producers: head of set i into head[i & 1]
    for (int i = 0; i < nloc; ++i) {
      const int buf = i & 1, set = blockIdx.x + i * gridDim.x;
      if (i >= 2) bar_sync(3 + buf, 512);  // the consumers have released this buffer
      if (b == 0) g.fill(ax[buf], set);
      bar_sync(5, 256);
      cd* hb = head + (size_t)buf * L;
      for (int j = b; j < L; j += 256) hb[j] = g(&ax[buf], set, set, j);
      bar_arrive(1 + buf, 512);
    }
    return;
  }
  // ---- consumers: the R/2 transforms of set i out of head[i & 1]
  const cd w0 = tw_load<-1>(twL, b);
  const cd w1 = tw_load<-1>(twL, (b & 15) * 16);
  const int base1 = fft_base16(b, 16);
  const cd wb = tw_load<-1>(twN, b);
  const int k0 = (b >> 4) + 16 * (b & 15);
  const CmEpi& ce = pe.inner;
  const double rk0 = (double)((long long)k0 << logR);
  const double al = ce.alpha * ce.lambda;
  for (int i = 0; i < nloc; ++i) {
    const int buf = i & 1, set = blockIdx.x + i * gridDim.x;
    bar_sync(1 + buf, 512);  // the head is complete
    const cd* hb = head + (size_t)buf * L;
    double* orow_set = ce.out + (size_t)(ce.idx ? ce.idx[set] : set) * ce.N;
    for (int pair = 0; pair < npairs; ++pair) {
      const int s = 2 * pair;
      const cd wbs = tw_load<-1>(twN, (b * s) & (N - 1));
      const cd cs0 = twC[s], cs1 = twC[s + 1];
      cd x[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int j = b + S0 * t;
        const cd xv = hb[j];
        if (j == 0) {
          x[t] = mk(xv.x, xv.x);
          continue;
        }
        const cd xr = conj(hb[L - j]);
        const cd w1t = wbs * twA[pair * 16 + t];
        const cd w2t = w1t * (wb * twB[t]);
        const cd a = w1t * (xv + cs0 * xr);
        const cd bb = w2t * (xv + cs1 * xr);
        x[t] = mk(0.5 * (a.x - bb.y), 0.5 * (a.y + bb.x));
      }
      if (pair == npairs - 1 && i + 2 < nloc) bar_arrive(3 + buf, 512);  // last read of this head
      dif16_regs_pow<-1>(x, w0);
      store16_dft<-1>(sm, b, S0, x);
      bar_sync(6, 256);
      load16(sm, base1, 16, x);
      dif16_regs_pow<-1>(x, w1);
      store16_dft<-1>(sm, base1, 16, x);
      bar_sync(6, 256);
      load16(sm, 16 * b, 1, x);
      Dft<16, -1>::run(x);
      // epilogue: factored undamping, see cm_pruned4096_kernel
      const double f0 = fm_exp(-al * s), f1 = fm_exp(-al * (s + 1));
      double ut = ce.S0 * fm_exp(-ce.alpha * (-ce.b + ce.lambda * rk0)) * ce.ada / (M_PI * 3.0);
      double kt = 0.0, g0 = 0.0, g1 = 0.0;
      if (ce.is_put) {
        g0 = fm_exp(ce.lambda * s);
        g1 = fm_exp(ce.lambda * (s + 1));
        kt = ce.S0 * fm_exp(-ce.b + ce.lambda * rk0) * ce.discount;
      }
      double* orow = orow_set + ((size_t)k0 << logR) + s;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const cd Z = x[Dft<16, -1>::out(t)];
        double v0 = Z.x * (ut * f0), v1 = Z.y * (ut * f1);
        if (ce.is_put) {
          v0 = v0 - ce.S0 + kt * g0;
          v1 = v1 - ce.S0 + kt * g1;
        }
        double* o = orow + ((size_t)(256 * t) << logR);
        if (pe.vec2) {
          *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
        } else {
          o[0] = v0;
          o[1] = v1;
        }
        ut *= pe.ustep;
        kt *= pe.kstep;
      }
      bar_sync(6, 256);  // the next transform overwrites the grid
    }
  }
}

// ---- integrateFFT functors (fft.hpp:64-82 genericIntegration)
struct IntGen {
  static constexpr bool has_cf = false;
  static size_t aux_bytes(int) { return 0; }
  const cd* fn;
  int N;
  double inputDx, outputMin;
  __device__ void prepare(void*, int, int) const {}
  // fn(x_j) inputDx e^{i outputMin inputDx j} / 3, Simpson weight 1 (ends) / 2 (even) / 4 (odd)
  __device__ cd operator()(const void*, int, int p, int j) const {
    const cd f = fn[(size_t)p * N + j];
    const cd ph = cis(outputMin * inputDx * j);
    const cd v = mk(f.x * inputDx, f.y * inputDx) * ph;
    const double wt = (j == 0 || j == N - 1) ? 1.0 : ((j & 1) ? 4.0 : 2.0);
    return mk(v.x / 3.0 * wt, v.y / 3.0 * wt);
  }
};
struct IntEpi {
  cd* out;
  int N;
  double outputMin, du, inputMin;
  __device__ void operator()(int p, int k, cd X) const {
    const double u = outputMin + k * du;
    out[(size_t)p * N + k] = X * cis(u * inputMin);
  }
};

}  // namespace

namespace {
using namespace fop;
// ---- FSTS on grids larger than the register-resident kernel of capi_fsts.cu covers (N > 4096) ----
// Same algorithm (OptionPricing.h:155-281) through the general FFT engine: the
// multiplier is evaluated once into HBM, every time step is forward transform (epilogue: multiply)
// and inverse transform (epilogue: real part, early-exercise max).  Throughput path for the sizes
// the reference's tests and BASELINE configs use is fsts_pair_kernel; this one removes the size
// limit of the drop-in.
struct FstsGrid {
  int N;
  double xmax, strike, dx;
  int payoff;
  __device__ double asset(int j) const { return strike * exp(-xmax + dx * j); }
  __device__ double pay(int j) const {
    const double sj = asset(j);
    return payoff == 1 ? (sj < strike ? strike - sj : 0.0) : (sj > strike ? sj - strike : 0.0);
  }
};
__global__ void __launch_bounds__(256) fsts_multiplier_kernel(int base, int tc, int np, const double* __restrict__ params,
                                                              FstsGrid gr, double r, double Tcf, double scale, int greek,
                                                              cd* __restrict__ mbuf) {
  __shared__ ModelConsts mc;
  const int p = blockIdx.y;
  if (threadIdx.x == 0) make_consts(base, tc, params + (size_t)p * np, mc);
  __syncthreads();
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= gr.N) return;
  const double vmax = M_PI / gr.dx, du = 2.0 * vmax / gr.N;
  double wk = du * k;
  if (wk > vmax) wk -= 2.0 * vmax;  // getUDomain, OptionPricing.h:19-23 (strict >)
  const cd u = mk(0.0, wk);
  CfPre pre;
  cf_prepare(base, tc, mc, u, r, pre);
  const cd phi = cexp(cf_log_at(tc, mc, pre, Tcf));
  const cd m = option_transform(greek, phi, u, r);
  mbuf[(size_t)p * gr.N + k] = mk(m.x * scale, m.y * scale);
}
__global__ void fsts_payoff_kernel(FstsGrid gr, int P, double* __restrict__ v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P * gr.N) return;
  v[i] = gr.pay((int)(i % gr.N));
}
struct FstsLoadReal {
  static constexpr bool has_cf = false;
  static size_t aux_bytes(int) { return 0; }
  const double* v;
  int N;
  __device__ void prepare(void*, int, int) const {}
  __device__ cd operator()(const void*, int, int p, int j) const { return mk(v[(size_t)p * N + j], 0.0); }
};
struct FstsMulEpi {
  const cd* m;
  cd* y;
  int N;
  __device__ void operator()(int p, int k, cd X) const {
    const size_t i = (size_t)p * N + k;
    y[i] = X * m[i];
  }
};
struct FstsRealEpi {
  double* v;
  FstsGrid gr;
  int american, kind_div;  // kind_div: 0 none, 1 / asset (delta), 2 / asset^2 (gamma); last step only
  __device__ void operator()(int p, int j, cd X) const {
    double x = X.x;
    if (american) x = fmax(x, gr.pay(j));
    if (kind_div) {
      const double sj = gr.asset(j);
      x = kind_div == 1 ? x / sj : x / (sj * sj);
    }
    v[(size_t)p * gr.N + j] = x;
  }
};

}  // namespace
namespace fop {
// device buffers only; d_out [P][N] receives the result
int fsts_large_run(fop_handle_t h, Arena& ar, int base, int tc, int np, const double* d_params, int P, int N,
                   double xmax, double strike, double r, double Tcf, double disc, int steps, int american,
                   int payoff, int kind, double* d_out) {
  const double dx = (2.0 * xmax) / ((double)N - 1.0);
  FstsGrid gr{N, xmax, strike, dx, payoff};
  const int greek = kind == 0 ? 0 : kind == 1 ? 1 : kind == 2 ? 2 : 3;
  // sets per chunk: multiplier + spectrum (32 B per point) within 2 GiB
  const int pcmax = (int)std::max<long long>(1, std::min<long long>(P, (2ll << 30) / ((long long)N * 32)));
  cd* mbuf = ar.alloc<cd>((size_t)pcmax * N);
  cd* ybuf = ar.alloc<cd>((size_t)pcmax * N);
  if (!mbuf || !ybuf) return fail(h, FOP_OOM, "arena exhausted (large-grid FSTS)");
  const size_t mark = h->arena_used;
  for (int p0 = 0; p0 < P; p0 += pcmax) {
    const int pc = std::min(pcmax, P - p0);
    double* v = d_out + (size_t)p0 * N;
    {
      ProfScope ps(h, "fsts_multiplier_kernel");
      fsts_multiplier_kernel<<<dim3((N + 255) / 256, pc), 256, 0, h->stream>>>(
          base, tc, np, d_params + (size_t)p0 * np, gr, r, Tcf, disc / (double)N, greek, mbuf);
    }
    FOP_LAUNCH_CHECK(h);
    fsts_payoff_kernel<<<(unsigned)(((size_t)pc * N + 255) / 256), 256, 0, h->stream>>>(gr, pc, v);
    FOP_LAUNCH_CHECK(h);
    for (int st = 0; st < steps; ++st) {
      const bool last = st == steps - 1;
      h->arena_used = mark;
      FOP_TRY((run_fft<-1>(h, ar, N, pc, FstsLoadReal{v, N}, FstsMulEpi{mbuf, ybuf, N})));
      h->arena_used = mark;
      FOP_TRY((run_fft<1>(h, ar, N, pc, LoadGen{ybuf, N},
                          FstsRealEpi{v, gr, american, last && (kind == 1 || kind == 2) ? kind : 0})));
    }
  }
  return FOP_OK;
}
size_t fsts_large_bytes(int N, int P) {
  const long long pc = std::max<long long>(1, std::min<long long>(P, (2ll << 30) / ((long long)N * 32)));
  return 2 * Arena::pad((size_t)pc * N * 16) + fft_scratch_bytes(N, (int)pc);
}
}  // namespace fop

extern "C" {

int fop_fft(fop_handle_t h, double* data, int batch, int N, int inverse) {
  if (!h) return FOP_INVALID_ARG;
  if (!data || batch < 0 || N < 2 || (N & (N - 1)) || N > (1 << 20))
    return fail(h, FOP_INVALID_ARG, "fop_fft: need data, batch >= 0, N a power of two in [2, 2^20]");
  if (batch == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const bool host = !is_device_ptr(data);
  const size_t bytes = (size_t)batch * N * 16;
  Arena ar(h);
  FOP_TRY(ar.reserve((host ? Arena::pad(bytes) : 0) + fft_scratch_bytes(N, batch)));
  cd* d = reinterpret_cast<cd*>(data);
  if (host) {
    d = ar.alloc<cd>((size_t)batch * N);
    FOP_CUDA(h, cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  if (N < 16) {
    tiny_dft_kernel<<<(batch + 127) / 128, 128, 0, h->stream>>>(d, batch, N, inverse ? 1 : -1);
    FOP_LAUNCH_CHECK(h);
  } else {
    LoadGen g{d, N};
    StoreEpi e{d, N};
    if (inverse) FOP_TRY((run_fft<1>(h, ar, N, batch, g, e)));
    else FOP_TRY((run_fft<-1>(h, ar, N, batch, g, e)));
  }
  if (host) {
    FOP_CUDA(h, cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}

int fop_integrate_fft(fop_handle_t h, const double* fn_samples, int batch, int N, double outputMin,
                      double inputMin, double inputMax, int inverse, double* out) {
  if (!h) return FOP_INVALID_ARG;
  if (!fn_samples || !out || batch < 0 || N < 16 || (N & (N - 1)) || N > (1 << 20) ||
      !(inputMax > inputMin))
    return fail(h, FOP_INVALID_ARG,
                "fop_integrate_fft: need samples, out, batch >= 0, N a power of two in [16, 2^20], "
                "inputMax > inputMin");
  if (batch == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const bool out_host = !is_device_ptr(out);
  const size_t count = (size_t)batch * N * 2;
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(fn_samples, count * 8) + (out_host ? Arena::pad(count * 8) : 0) +
                     fft_scratch_bytes(N, batch)));
  const double* d_fn = nullptr;
  FOP_TRY(stage_in(h, ar, fn_samples, count, &d_fn));
  double* d_out = out_host ? ar.alloc<double>(count) : out;
  const double range = inputMax - inputMin;
  const double inputDx = range / (N - 1);
  const double outputMax = outputMin + (N - 1) * 2.0 * M_PI / range;
  const double du = (outputMax - outputMin) / N;
  IntGen g{reinterpret_cast<const cd*>(d_fn), N, inputDx, outputMin};
  IntEpi e{reinterpret_cast<cd*>(d_out), N, outputMin, du, inputMin};
  if (inverse) FOP_TRY((run_fft<1>(h, ar, N, batch, g, e)));
  else FOP_TRY((run_fft<-1>(h, ar, N, batch, g, e)));
  if (out_host) {
    FOP_CUDA(h, cudaMemcpyAsync(out, d_out, count * 8, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}

int fop_carr_madan(fop_handle_t h, fop_model_t model, const double* params, int P, int N,
                   double ada, double alpha, double S0, double r, double T, int is_put,
                   double* out) {
  if (!h) return FOP_INVALID_ARG;
  const int np = num_params(model.base, model.time_change);
  if (np == 0) return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if (!params || !out) return fail(h, FOP_INVALID_ARG, "null pointer");
  if (P < 0 || N < 16 || (N & (N - 1)) || N > (1 << 20))
    return fail(h, FOP_INVALID_ARG, "fop_carr_madan: N must be a power of two in [16, 2^20]");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const bool out_host = !is_device_ptr(out);
  // host-bound outputs are produced in chunks of <= 1 GiB of device staging
  const long long chunkP = out_host ? std::max<long long>(1, std::min<long long>(P, (1ll << 30) / ((long long)N * 8))) : P;
  const bool no_skip = std::getenv("FOP_CM_NO_SKIP") != nullptr;
  // long grids of a Levy model: route the sets whose non-zero head fits a single-CTA transform
  // through the input-pruned path (FOP_CM_NO_PRUNE=1: A/B runs)
  const bool try_prune = N > 8192 && model.time_change == FOP_TC_NONE && !no_skip &&
                         !std::getenv("FOP_CM_NO_PRUNE");
  // the pruned path needs head buffers (sized from the real work, <= 512 MiB), the transform scratch
  // of its 2^14 / 2^15-point heads, and two routing index arrays -- nothing when it is not taken
  size_t prune_bytes = 0;
  if (try_prune) {
    const long long Lmax = std::min<long long>(32768, N / 2);
    const long long head = std::min<long long>(chunkP * Lmax * 16, 512ll << 20);
    const long long tb = std::max<long long>(1, head / (Lmax * 16));
    prune_bytes = Arena::pad((size_t)head) +
                  fft_scratch_bytes((int)Lmax, (int)std::min<long long>(tb * (N / Lmax) / 2 + 1, 1 << 30)) +
                  2 * Arena::pad((size_t)chunkP * 4) + Arena::pad((size_t)chunkP * sizeof(CmAux));
  }
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(params, (size_t)P * np * 8) +
                     (out_host ? Arena::pad((size_t)chunkP * N * 8) : 0) +
                     fft_scratch_bytes(N, (int)chunkP) + prune_bytes));
  const double* d_params = nullptr;
  FOP_TRY(stage_in(h, ar, params, (size_t)P * np, &d_params));
  double* d_tmp = out_host ? ar.alloc<double>((size_t)chunkP * N) : nullptr;
  const double b = M_PI / ada;        // getMaxK,  OptionPricing.h:31-33
  const double lambda = 2.0 * b / N;  // getLambda, :40-42
  const double discount = std::exp(-r * T);
  const size_t scratch_mark = h->arena_used;
  for (long long p0 = 0; p0 < P; p0 += chunkP) {
    const int pc = (int)std::min<long long>(chunkP, P - p0);
    h->arena_used = scratch_mark;  // scratch is re-used by every chunk
    CmGen g{model.base, model.time_change, np, d_params + (size_t)p0 * np, r, T, ada, alpha, discount, b,
            no_skip, nullptr};
    CmEpi e{out_host ? d_tmp : out + (size_t)p0 * N, N, S0, alpha, b, lambda, ada, discount, is_put, nullptr};
    if (!try_prune) {
      FOP_TRY((run_fft<-1>(h, ar, N, pc, g, e)));
    } else {
      // 1. cut-off index of every set (device), lists per head length (host)
      int* d_jc = ar.alloc<int>(pc);
      int* d_idx = ar.alloc<int>(pc);
      if (!d_jc || !d_idx) return fail(h, FOP_OOM, "arena exhausted (Carr-Madan routing)");
      CmAux* d_aux = ar.alloc<CmAux>(pc);
      if (!d_aux) return fail(h, FOP_OOM, "arena exhausted (Carr-Madan set constants)");
      cm_cutoff_kernel<<<(pc + 127) / 128, 128, 0, h->stream>>>(g, pc, N, d_jc, d_aux);
      g.aux_all = d_aux;  // the kernels below load the per-set constants instead of re-forming them
      FOP_LAUNCH_CHECK(h);
      std::vector<int> jc(pc);
      FOP_CUDA(h, cudaMemcpyAsync(jc.data(), d_jc, sizeof(int) * pc, cudaMemcpyDeviceToHost, h->stream));
      FOP_CUDA(h, cudaStreamSynchronize(h->stream));
      // head lengths 2^12 .. 2^15 (at most N / 2); everything else takes the plain four-step path
      constexpr int NG = 4;
      std::vector<int> lists[NG + 1];
      for (int i = 0; i < pc; ++i) {
        int gsel = NG;
        for (int q = 0; q < NG; ++q)
          if (jc[i] <= (4096 << q) && (4096 << q) <= N / 2) { gsel = q; break; }
        lists[gsel].push_back(i);
      }
      std::vector<int> order;
      for (auto& l : lists) order.insert(order.end(), l.begin(), l.end());
      FOP_CUDA(h, cudaMemcpyAsync(d_idx, order.data(), sizeof(int) * pc, cudaMemcpyHostToDevice, h->stream));
      FOP_CUDA(h, cudaStreamSynchronize(h->stream));  // `order` goes out of scope below
      const double2* twN = nullptr;
      FOP_TRY(get_twiddles(h, N, &twN));
      const size_t mark2 = h->arena_used;
      int off = 0;
      for (int grp = 0; grp <= NG; ++grp) {
        const int cnt = (int)lists[grp].size();
        if (cnt == 0) continue;
        h->arena_used = mark2;
        CmGen gg = g;
        CmEpi ee = e;
        if (grp == NG) {
          gg.idx = d_idx + off;
          ee.idx = d_idx + off;
          FOP_TRY((run_fft<-1>(h, ar, N, cnt, gg, ee)));
        } else {
          const int logL = 12 + grp, L = 1 << logL, logR = ilog2(N) - logL;
          // head buffers for up to `tb` sets at a time (<= 512 MiB)
          const int tb = (int)std::min<long long>(cnt, (512ll << 20) / ((long long)L * 16));
          cd* xbuf = ar.alloc<cd>((size_t)tb * L);
          if (!xbuf) return fail(h, FOP_OOM, "arena exhausted (Carr-Madan head buffer)");
          const size_t mark3 = h->arena_used;
          for (int t0 = 0; t0 < cnt; t0 += tb) {
            const int tc = std::min(tb, cnt - t0);
            gg.idx = d_idx + off + t0;
            ee.idx = d_idx + off + t0;
            // FOP_CM_PRUNED=set | ws: the set-per-CTA experiments (head in shared memory; phases in turn
            // or warp-specialised).  Measured on config 4 (profiles/r2_cm_pruned_variants.json): 8.9 /
            // 11.0 ms against 8.5 ms for head kernel + cm_pruned4096_kernel, which stays the default.
            const char* pv = std::getenv("FOP_CM_PRUNED");
            if (logL == 12 && logR <= 6 && pv && (!std::strcmp(pv, "set") || !std::strcmp(pv, "ws")) &&
                !std::getenv("FOP_CM_GENERIC_PRUNED")) {
              // set-per-CTA kernels: head evaluated into shared memory, no head buffer round trip
              const double2* twL = nullptr;
              FOP_TRY(get_twiddles(h, L, &twL));
              PrunedEpi pe{ee, logR, reinterpret_cast<uintptr_t>(ee.out) % 16 == 0 ? 1 : 0,
                           std::exp(-alpha * lambda * 256.0 * (double)(1 << logR)),
                           std::exp(lambda * 256.0 * (double)(1 << logR))};
              const bool ws = !std::strcmp(pv, "ws");
              const size_t smem = (ws ? (2 * (size_t)L + (size_t)fft_padded_size(L)) : ((size_t)L + 2 * (size_t)fft_padded_size(L))) * sizeof(cd) +
                                  2 * sizeof(CmAux) + ((size_t)(16 + 2) * (1 << (logR - 1)) + 16) * sizeof(cd);
              auto kern = ws ? cm_pruned4096_ws_kernel : cm_pruned4096_set_kernel;
              FOP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              {
                ProfScope ps(h, ws ? "carr_madan_pruned_ws_kernel" : "carr_madan_pruned_set_kernel");
                kern<<<std::min(tc, h->sm_count), 512, smem, h->stream>>>(gg, pe, tc, N, twN, twL);
              }
              FOP_LAUNCH_CHECK(h);
              continue;
            }
            {
              ProfScope ps(h, "carr_madan_head_kernel");
              cm_head_kernel<<<tc, 256, 0, h->stream>>>(gg, L, xbuf);
            }
            FOP_LAUNCH_CHECK(h);
            PrunedGen pg{xbuf, twN, L, logR, N};
            PrunedEpi pe{ee, logR, reinterpret_cast<uintptr_t>(ee.out) % 16 == 0 ? 1 : 0,
                         std::exp(-alpha * lambda * 256.0 * (double)(1 << logR)),
                         std::exp(lambda * 256.0 * (double)(1 << logR))};
            const int batch = tc << (logR - 1);  // two residue classes per transform
            if (logL == 12 && !std::getenv("FOP_CM_GENERIC_PRUNED")) {
              const double2* twL = nullptr;
              FOP_TRY(get_twiddles(h, L, &twL));
              const size_t smem = ((size_t)fft_padded_size(L) + (size_t)(16 + 2) * (1 << (logR - 1)) + 16) * sizeof(cd);
              FOP_CUDA(h, cudaFuncSetAttribute(cm_pruned4096_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              {
                ProfScope ps(h, "carr_madan_pruned_kernel");
                cm_pruned4096_kernel<<<std::min(batch, h->sm_count * 16), 256, smem, h->stream>>>(pg, pe, batch, twL);
              }
              FOP_LAUNCH_CHECK(h);
            } else if (logL <= 13) {  // single-CTA transform, length fixed at compile time
              const double2* twL = nullptr;
              FOP_TRY(get_twiddles(h, L, &twL));
              auto kern = logL == 12 ? fft_single_kernel<-1, PrunedGen, PrunedEpi, 12>
                                     : fft_single_kernel<-1, PrunedGen, PrunedEpi, 13>;
              const size_t smem = fft_single_smem(logL);
              FOP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              const int tpbk = fft_single_tpb(logL);
              const int grid = std::min((batch + tpbk - 1) / tpbk, h->sm_count * 16);
              {
                ProfScope ps(h, "carr_madan_pruned_kernel");
                kern<<<grid, fft_single_threads(logL), smem, h->stream>>>(logL, batch, twL, pg, pe);
              }
              FOP_LAUNCH_CHECK(h);
            } else {  // 2^14, 2^15: the pruned transforms themselves go through the four-step engine
              h->arena_used = mark3;
              FOP_TRY((run_fft<-1>(h, ar, L, batch, pg, pe)));
            }
          }
        }
        off += cnt;
      }
    }
    if (out_host)
      FOP_CUDA(h, cudaMemcpyAsync(out + (size_t)p0 * N, d_tmp, (size_t)pc * N * 8,
                                  cudaMemcpyDeviceToHost, h->stream));
  }
  if (out_host) FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  return FOP_OK;
}

}  // extern "C"

// ================================================================================================
// Caller-sampled entry points.  The reference takes ANY callable as characteristic function,
// payoff and asset map (FSTSGeneric OptionPricing.h:155-163, CarrMadanGeneric :582-590); a device
// kernel cannot call a host lambda, but it does not need to: the callable is only ever evaluated
// on a grid the library defines, so the host samples it there (fop_fsts_u_grid /
// fop_carr_madan_u_grid give the points) and everything after the evaluation -- option transform,
// damping, Simpson weights, transforms, time stepping, early exercise, output map -- runs here.
// ================================================================================================
namespace {
struct CmSampleGen {
  static constexpr bool has_cf = false;
  static size_t aux_bytes(int) { return 0; }
  const cd* phi;  // [P][N]: cf(alpha + 1 + i j ada)
  int N;
  double ada, alpha, discount, b;
  __device__ void prepare(void*, int, int) const {}
  // same arithmetic as CmGen::operator() after the CF evaluation (OptionPricing.h:594-599)
  __device__ cd operator()(const void*, int, int p, int j) const {
    const double w = j * ada;
    const cd den = mk(alpha * alpha + alpha - w * w, (2.0 * alpha + 1.0) * w);
    const cd aug = cdiv(phi[(size_t)p * N + j], den);
    const cd ph = cis(b * j * ada);
    const double wt = j == 0 ? 1.0 : ((j & 1) ? 4.0 : 2.0);
    const cd v = aug * ph;
    const double sc = discount * wt;
    return mk(v.x * sc, v.y * sc);
  }
};
// m_k = enhCF(phi_k, i w_k) * scale from sampled phi (FFT order, getUDomain)
__global__ void fsts_multiplier_samples_kernel(const cd* __restrict__ phi, int N, double dx, double r,
                                               double scale, int greek, cd* __restrict__ mbuf) {
  const int p = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= N) return;
  const double vmax = M_PI / dx, du = 2.0 * vmax / N;
  double wk = du * k;
  if (wk > vmax) wk -= 2.0 * vmax;  // getUDomain, OptionPricing.h:19-23 (strict >)
  const cd m = option_transform(greek, phi[(size_t)p * N + k], mk(0.0, wk), r);
  mbuf[(size_t)p * N + k] = mk(m.x * scale, m.y * scale);
}
__global__ void fsts_payoff_samples_kernel(const double* __restrict__ pay, int per_set, int N, int P,
                                           double* __restrict__ v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P * N) return;
  v[i] = pay[per_set ? i : i % N];
}
struct FstsRealSamplesEpi {
  double* v;
  const double* pay;    // [N] or [P][N]
  const double* asset;  // [N] or nullptr
  int N, per_set, american, kind_div;
  __device__ void operator()(int p, int j, cd X) const {
    double x = X.x;
    if (american) x = fmax(x, pay[per_set ? (size_t)p * N + j : (size_t)j]);
    if (kind_div) {
      const double sj = asset[j];
      x = kind_div == 1 ? x / sj : x / (sj * sj);
    }
    v[(size_t)p * N + j] = x;
  }
};
}  // namespace

extern "C" {

int fop_carr_madan_u_grid(int N, double ada, double alpha, double* u_out2) {
  if (!u_out2 || N < 1) return FOP_INVALID_ARG;
  for (int j = 0; j < N; ++j) {
    u_out2[2 * j] = alpha + 1.0;
    u_out2[2 * j + 1] = j * ada;
  }
  return FOP_OK;
}

int fop_carr_madan_cf_samples(fop_handle_t h, const double* cf_samples, int P, int N, double ada,
                              double alpha, double S0, double r, double T, int is_put, double* out) {
  if (!h) return FOP_INVALID_ARG;
  if (!cf_samples || !out) return fail(h, FOP_INVALID_ARG, "null pointer");
  if (P < 0 || N < 16 || (N & (N - 1)) || N > (1 << 20))
    return fail(h, FOP_INVALID_ARG, "fop_carr_madan_cf_samples: N must be a power of two in [16, 2^20]");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const bool out_host = !is_device_ptr(out);
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(cf_samples, (size_t)P * N * 16) + (out_host ? Arena::pad((size_t)P * N * 8) : 0) +
                     fft_scratch_bytes(N, P)));
  const double* d_phi = nullptr;
  FOP_TRY(stage_in(h, ar, cf_samples, (size_t)P * N * 2, &d_phi));
  double* d_out = out_host ? ar.alloc<double>((size_t)P * N) : out;
  const double b = M_PI / ada, lambda = 2.0 * b / N, discount = std::exp(-r * T);
  CmSampleGen g{reinterpret_cast<const cd*>(d_phi), N, ada, alpha, discount, b};
  CmEpi e{d_out, N, S0, alpha, b, lambda, ada, discount, is_put, nullptr};
  FOP_TRY((run_fft<-1>(h, ar, N, P, g, e)));
  if (out_host) {
    FOP_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)P * N * 8, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}

int fop_fsts_u_grid(int N, double xmax, double* w_out) {
  if (!w_out || N < 2) return FOP_INVALID_ARG;
  const double dx = (2.0 * xmax) / ((double)N - 1.0), vmax = M_PI / dx, du = 2.0 * vmax / N;
  for (int k = 0; k < N; ++k) {
    double wk = du * k;
    if (wk > vmax) wk -= 2.0 * vmax;
    w_out[k] = wk;
  }
  return FOP_OK;
}

int fop_fsts_samples(fop_handle_t h, const double* cf_samples, const double* payoff, int payoff_per_set,
                     const double* asset, int P, int N, double xmax, double r, double T, int steps,
                     int american, int kind, double* out) {
  if (!h) return FOP_INVALID_ARG;
  if (!cf_samples || !payoff || !out) return fail(h, FOP_INVALID_ARG, "null pointer");
  if (P < 0 || N < 16 || (N & (N - 1)) || N > (1 << 20) || steps < 1 || kind < 0 || kind > 3)
    return fail(h, FOP_INVALID_ARG, "fop_fsts_samples: N power of two in [16, 2^20], steps >= 1");
  if ((kind == FOP_FSTS_DELTA || kind == FOP_FSTS_GAMMA) && !asset)
    return fail(h, FOP_INVALID_ARG, "fop_fsts_samples: delta / gamma need the asset grid");
  if ((steps > 1 || american) && kind != FOP_FSTS_PRICE)
    return fail(h, FOP_UNSUPPORTED, "fop_fsts_samples: time stepping / early exercise needs kind = price");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const bool out_host = !is_device_ptr(out);
  const size_t npay = payoff_per_set ? (size_t)P * N : (size_t)N;
  const long long pcmax = std::max<long long>(1, std::min<long long>(P, (2ll << 30) / ((long long)N * 32)));
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(cf_samples, (size_t)P * N * 16) + stage_bytes(payoff, npay * 8) +
                     (asset ? stage_bytes(asset, (size_t)N * 8) : 0) +
                     (out_host ? Arena::pad((size_t)P * N * 8) : 0) +
                     2 * Arena::pad((size_t)pcmax * N * 16) + fft_scratch_bytes(N, (int)pcmax)));
  const double *d_phi = nullptr, *d_pay = nullptr, *d_asset = nullptr;
  FOP_TRY(stage_in(h, ar, cf_samples, (size_t)P * N * 2, &d_phi));
  FOP_TRY(stage_in(h, ar, payoff, npay, &d_pay));
  if (asset) FOP_TRY(stage_in(h, ar, asset, (size_t)N, &d_asset));
  double* d_out = out_host ? ar.alloc<double>((size_t)P * N) : out;
  cd* mbuf = ar.alloc<cd>((size_t)pcmax * N);
  cd* ybuf = ar.alloc<cd>((size_t)pcmax * N);
  if (!mbuf || !ybuf || !d_out) return fail(h, FOP_OOM, "arena exhausted (fop_fsts_samples)");
  const double dx = (2.0 * xmax) / ((double)N - 1.0);
  const double Tcf = T / steps, disc = std::exp(-r * Tcf);
  const int greek = kind == 0 ? 0 : kind == 1 ? 1 : kind == 2 ? 2 : 3;
  const size_t mark = h->arena_used;
  for (long long p0 = 0; p0 < P; p0 += pcmax) {
    const int pc = (int)std::min<long long>(pcmax, P - p0);
    double* v = d_out + (size_t)p0 * N;
    const double* pay0 = d_pay + (payoff_per_set ? (size_t)p0 * N : 0);
    {
      ProfScope ps(h, "fsts_multiplier_kernel");
      fsts_multiplier_samples_kernel<<<dim3((N + 255) / 256, pc), 256, 0, h->stream>>>(
          reinterpret_cast<const cd*>(d_phi) + (size_t)p0 * N, N, dx, r, disc / (double)N, greek, mbuf);
    }
    FOP_LAUNCH_CHECK(h);
    fsts_payoff_samples_kernel<<<(unsigned)(((size_t)pc * N + 255) / 256), 256, 0, h->stream>>>(
        pay0, payoff_per_set, N, pc, v);
    FOP_LAUNCH_CHECK(h);
    for (int st = 0; st < steps; ++st) {
      const bool last = st == steps - 1;
      h->arena_used = mark;
      FOP_TRY((run_fft<-1>(h, ar, N, pc, FstsLoadReal{v, N}, FstsMulEpi{mbuf, ybuf, N})));
      h->arena_used = mark;
      FOP_TRY((run_fft<1>(h, ar, N, pc, LoadGen{ybuf, N},
                          FstsRealSamplesEpi{v, pay0, d_asset, N, payoff_per_set, american ? 1 : 0,
                                             last && (kind == 1 || kind == 2) ? kind : 0})));
    }
  }
  if (out_host) {
    FOP_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)P * N * 8, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}

}  // extern "C"
