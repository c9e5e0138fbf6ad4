// C ABI: fop_calibrate_cos -- population calibration of the PRICE-space objective of fop_cos_loss
// (BASELINE.json config 5: the parameter sweep turned into a search).
//
// Same cuckoo search as fop_calibrate_cf (capi_calibrate.cu: the semantics of the reference's
// cuckoo::optimize call, generateCharts.cpp:102-118 -- Levy flights towards the best nest,
// abandonment by a walk between two other nests, greedy acceptance, counter-based random numbers),
// but the objective of a candidate is the weighted squared error of its COS prices against the market
// surface, so one objective evaluation is a whole FangOostCallPrice surface (OptionPricing.h:362-382).
// B200 mapping: `restarts` independent searches of `nests` nests form ONE population of restarts x
// nests parameter sets; every half-iteration proposes candidates for the whole population (one small
// kernel), prices them with the batched COS path (coefficient kernel + tensor-core contraction +
// fused loss: the same kernels and the same bits as fop_cos_loss), and accepts greedily (one small
// kernel).  Nothing returns to the host inside the loop except, every 16 iterations, the best
// objective value for the tolerance test.
#include <algorithm>
#include <cmath>
#include <vector>

#include "cf_models.cuh"
#include "handle.cuh"

namespace {
using namespace fop;

constexpr int CKP_MAX_NP = 9;

__device__ __forceinline__ unsigned long long ckp_splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct CkpRng {  // one stream per (seed, restart, iteration, phase, nest, coordinate), as in capi_calibrate.cu
  unsigned long long key, ctr;
  __device__ CkpRng(unsigned long long seed, int restart, int it, int phase, int nest, int dim) {
    key = ckp_splitmix64(seed ^ ckp_splitmix64(((unsigned long long)restart << 32) | (unsigned)it));
    key = ckp_splitmix64(key ^ (((unsigned long long)phase << 48) | ((unsigned long long)nest << 16) | (unsigned)dim));
    ctr = 0;
  }
  __device__ double uniform() {
    const unsigned long long z = ckp_splitmix64(key + (ctr++) * 0xD1342543DE82EF95ull);
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double normal() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
};

struct CkpState {
  int np, nests, restarts;
  unsigned long long seed;
  const double* lo;
  const double* hi;
  double* pos;    // [restarts][nests][np]
  double* val;    // [restarts][nests]
  double* cand;   // [restarts][nests][np]
  double* cval;   // [restarts][nests]
  int* best;      // [restarts]
};

// phase 0: initial nests; 1: Levy flight relative to the restart's best nest; 2: abandonment walk
__global__ void ckp_propose_kernel(CkpState s, int phase, int it) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = s.restarts * s.nests * s.np;
  if (idx >= total) return;
  const int d = idx % s.np, i = (idx / s.np) % s.nests, rs = idx / (s.np * s.nests);
  const double lo = s.lo[d], hi = s.hi[d];
  const double* pos = s.pos + (size_t)rs * s.nests * s.np;
  double x;
  if (phase == 0) {
    CkpRng g(s.seed, rs, 0, 0, i, d);
    x = lo + (hi - lo) * g.uniform();
  } else if (phase == 1) {
    const double sigma_u = 0.6965745025576968;  // Mantegna, beta = 1.5
    CkpRng g(s.seed, rs, it + 1, 1, i, d);
    const double un = g.normal() * sigma_u, vn = g.normal();
    const double step = un / pow(fabs(vn), 1.0 / 1.5);
    x = pos[i * s.np + d];
    x += 0.01 * step * (x - pos[s.best[rs] * s.np + d]) * g.normal();
  } else {
    CkpRng gi(s.seed, rs, it + 1, 2, i, CKP_MAX_NP);
    const int j1 = (int)(gi.uniform() * s.nests) % s.nests, j2 = (int)(gi.uniform() * s.nests) % s.nests;
    const double scale = gi.uniform();
    CkpRng g(s.seed, rs, it + 1, 2, i, d);
    x = pos[i * s.np + d];
    if (g.uniform() < 0.25) x += scale * (pos[j1 * s.np + d] - pos[j2 * s.np + d]);
  }
  s.cand[idx] = fmin(fmax(x, lo), hi);
}
// greedy acceptance (NaN never accepted); init: take everything
__global__ void ckp_accept_kernel(CkpState s, int init) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.restarts * s.nests) return;
  const bool take = init ? true : s.cval[i] < s.val[i];
  if (take) {
    for (int d = 0; d < s.np; ++d) s.pos[(size_t)i * s.np + d] = s.cand[(size_t)i * s.np + d];
    s.val[i] = s.cval[i];
  }
}
// best nest of every restart (ties -> lowest index, NaN never best unless all are)
__global__ void ckp_best_kernel(CkpState s, double* bestval) {
  const int rs = blockIdx.x * blockDim.x + threadIdx.x;
  if (rs >= s.restarts) return;
  const double* v = s.val + (size_t)rs * s.nests;
  int b = 0;
  for (int i = 1; i < s.nests; ++i)
    if (v[i] < v[b] || (v[b] != v[b] && v[i] == v[i])) b = i;
  s.best[rs] = b;
  bestval[rs] = v[b];
}
__global__ void ckp_output_kernel(CkpState s, double* params_out, double* fnval_out) {
  const int rs = blockIdx.x * blockDim.x + threadIdx.x;
  if (rs >= s.restarts) return;
  const int b = s.best[rs];
  for (int d = 0; d < s.np; ++d) params_out[(size_t)rs * s.np + d] = s.pos[((size_t)rs * s.nests + b) * s.np + d];
  fnval_out[rs] = s.val[(size_t)rs * s.nests + b];
}

}  // namespace

extern "C" int fop_calibrate_cos(fop_handle_t h, fop_model_t model, const double* lower, const double* upper,
                                 const double* K_desc, int J, double S0, double r, const double* T, int M,
                                 int numU, const double* mkt, const double* w, int nests, int iterations,
                                 double tol, unsigned long long seed, int restarts, double* params_out,
                                 double* fnval_out, int* iters_out, int* best_restart_out) {
  if (!h) return FOP_INVALID_ARG;
  const int np = num_params(model.base, model.time_change);
  if (np == 0) return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if (!lower || !upper || !K_desc || !T || !mkt || !params_out || !fnval_out)
    return fail(h, FOP_INVALID_ARG, "null pointer");
  if (nests < 2 || nests > 4096 || iterations < 0 || restarts < 1 || J < 2 || M < 1)
    return fail(h, FOP_INVALID_ARG, "fop_calibrate_cos: need 2 <= nests <= 4096, iterations >= 0, restarts >= 1");
  FOP_CUDA(h, cudaSetDevice(h->device));
  double lo_h[16], hi_h[16];
  FOP_CUDA(h, cudaMemcpy(lo_h, lower, (size_t)np * 8, cudaMemcpyDefault));
  FOP_CUDA(h, cudaMemcpy(hi_h, upper, (size_t)np * 8, cudaMemcpyDefault));
  for (int i = 0; i < np; ++i)
    if (!(hi_h[i] >= lo_h[i])) return fail(h, FOP_INVALID_ARG, "bound %d: upper < lower", i);
  const size_t pop = (size_t)restarts * nests;
  if (pop > (1u << 30) / (size_t)np) return fail(h, FOP_INVALID_ARG, "population too large");
  // search state: outside the per-call arena (fop_cos_loss below resets it on every call)
  const size_t nd = 2 * (size_t)np + 2 * pop * np + 2 * pop + restarts + (size_t)M * J * 2;
  double* d_state = nullptr;
  int* d_best = nullptr;
  FOP_CUDA(h, cudaMalloc(&d_state, nd * 8));
  if (cudaMalloc(&d_best, (size_t)restarts * 4) != cudaSuccess) {
    cudaFree(d_state);
    return fail(h, FOP_OOM, "fop_calibrate_cos: out of memory");
  }
  auto cleanup = [&]() { cudaFree(d_state); cudaFree(d_best); };
#define CKP_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) {                                   \
      cleanup();                                                \
      return fail(h, FOP_CUDA_ERROR, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    }                                                           \
  } while (0)
#define CKP_TRY(expr)            \
  do {                           \
    int st__ = (expr);           \
    if (st__ != FOP_OK) {        \
      cleanup();                 \
      return st__;               \
    }                            \
  } while (0)
  double* p = d_state;
  double* d_lo = p;  p += np;
  double* d_hi = p;  p += np;
  CkpState s;
  s.np = np;  s.nests = nests;  s.restarts = restarts;  s.seed = seed;
  s.lo = d_lo;  s.hi = d_hi;
  s.pos = p;  p += pop * np;
  s.cand = p;  p += pop * np;
  s.val = p;  p += pop;
  s.cval = p;  p += pop;
  double* d_bestval = p;  p += restarts;
  double* d_mkt = p;  p += (size_t)M * J;
  double* d_w = w ? p : nullptr;
  s.best = d_best;
  CKP_CUDA(cudaMemcpyAsync(d_lo, lo_h, (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
  CKP_CUDA(cudaMemcpyAsync(d_hi, hi_h, (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
  CKP_CUDA(cudaMemcpyAsync(d_mkt, mkt, (size_t)M * J * 8, cudaMemcpyDefault, h->stream));
  if (w) CKP_CUDA(cudaMemcpyAsync(d_w, w, (size_t)M * J * 8, cudaMemcpyDefault, h->stream));
  CKP_CUDA(cudaMemsetAsync(d_best, 0, (size_t)restarts * 4, h->stream));
  // strike / maturity descriptors on the host once (fop_cos_loss reads host descriptors directly)
  std::vector<double> Kh(J), Th(M);
  CKP_CUDA(cudaMemcpy(Kh.data(), K_desc, (size_t)J * 8, cudaMemcpyDefault));
  CKP_CUDA(cudaMemcpy(Th.data(), T, (size_t)M * 8, cudaMemcpyDefault));

  const int tpb = 256;
  const unsigned gp = (unsigned)((pop * np + tpb - 1) / tpb), ga = (unsigned)((pop + tpb - 1) / tpb),
                 gr = (unsigned)((restarts + tpb - 1) / tpb);
  auto launch_check = [&]() -> int {
    h->launches++;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FOP_OK : fail(h, FOP_CUDA_ERROR, "kernel launch failed: %s", cudaGetErrorString(e));
  };
  auto evaluate = [&]() -> int {  // cval = price-space objective of every candidate (device to device)
    return fop_cos_loss(h, model, s.cand, (int)pop, Kh.data(), J, S0, r, Th.data(), M, numU, d_mkt, d_w, s.cval,
                        nullptr);
  };
  auto half_step = [&](int phase, int it, int init) -> int {
    ckp_propose_kernel<<<gp, tpb, 0, h->stream>>>(s, phase, it);
    FOP_TRY(launch_check());
    FOP_TRY(evaluate());
    ckp_accept_kernel<<<ga, tpb, 0, h->stream>>>(s, init);
    return launch_check();
  };
  CKP_TRY(half_step(0, 0, 1));
  ckp_best_kernel<<<gr, tpb, 0, h->stream>>>(s, d_bestval);
  CKP_TRY(launch_check());
  int it = 0;
  std::vector<double> bv(restarts);
  for (; it < iterations; ++it) {
    CKP_TRY(half_step(1, it, 0));
    CKP_TRY(half_step(2, it, 0));
    ckp_best_kernel<<<gr, tpb, 0, h->stream>>>(s, d_bestval);
    CKP_TRY(launch_check());
    if ((it & 15) == 15 && tol > 0.0) {  // tolerance test: any restart below tol ends the search
      CKP_CUDA(cudaMemcpyAsync(bv.data(), d_bestval, (size_t)restarts * 8, cudaMemcpyDeviceToHost, h->stream));
      CKP_CUDA(cudaStreamSynchronize(h->stream));
      if (*std::min_element(bv.begin(), bv.end()) < tol) {
        ++it;
        break;
      }
    }
  }
  // outputs (host or device)
  double* d_pout = s.cand;   // candidates are dead now: reuse as staging
  double* d_fout = s.cval;
  ckp_output_kernel<<<gr, tpb, 0, h->stream>>>(s, d_pout, d_fout);
  CKP_TRY(launch_check());
  CKP_CUDA(cudaMemcpyAsync(params_out, d_pout, (size_t)restarts * np * 8, cudaMemcpyDefault, h->stream));
  CKP_CUDA(cudaMemcpyAsync(fnval_out, d_fout, (size_t)restarts * 8, cudaMemcpyDefault, h->stream));
  CKP_CUDA(cudaMemcpyAsync(bv.data(), d_fout, (size_t)restarts * 8, cudaMemcpyDeviceToHost, h->stream));
  CKP_CUDA(cudaStreamSynchronize(h->stream));
  if (iters_out) {
    std::vector<int> its(restarts, it);
    CKP_CUDA(cudaMemcpy(iters_out, its.data(), (size_t)restarts * 4, cudaMemcpyDefault));
  }
  if (best_restart_out) {
    int b = 0;
    for (int i = 1; i < restarts; ++i)
      if (bv[i] < bv[b] || (bv[b] != bv[b] && bv[i] == bv[i])) b = i;
    *best_restart_out = b;
  }
  cleanup();
  return FOP_OK;
#undef CKP_CUDA
#undef CKP_TRY
}
