// C ABI: fop_calibrate_cf -- on-device population search (cuckoo search) over the calibration
// objective of optioncal::getObjFn_arr (OptionCalibration.h:185-201).
//
// Replaces the reference's use of cuckoo::optimize(objFn, constraints, nestSize = 25, numSims =
// 1500, tol, seed) in generateCharts.cpp:115 (library danielhstahl/cuckoo_search: un-vendored and
// unpinned, .travis.yml:25 -- there is no reference code to be bit-compatible with; the algorithm
// below is the published one, Yang & Deb 2009 / Yang's reference implementation: Levy flights
// towards the best nest by Mantegna's algorithm, beta = 3/2, step 0.01; then abandonment of a
// fraction pa = 1/4 of the coordinates by a biased random walk between two other nests; greedy
// acceptance).  Same semantics for the caller: bounds, nest count, iteration cap, tolerance, seed
// in; best parameter vector and objective value out.
//
// B200 mapping: the whole search runs inside ONE kernel launch.  One CTA per independent restart
// (restarts >= the SM count fill the GPU; the host keeps the best), 8 warps per CTA, nests dealt
// round-robin to warps, a warp evaluates one objective with its lanes over u_l (warp_objfn_cf --
// the same code, hence the same bits, as fop_objfn_cf).  Random numbers are counter based
// (splitmix64 of seed / restart / iteration / phase / nest / coordinate), so a run is
// reproducible and independent of scheduling.  No host round trip per iteration.
#include <algorithm>
#include <cmath>
#include <vector>

#include "handle.cuh"
#include "objfn_common.cuh"

namespace {
using namespace fop;

constexpr int CK_WARPS = 8;
constexpr int CK_MAX_NESTS = 64;
constexpr int CK_MAX_NP = 9;

struct CuckooArgs {
  int base, tc, np, m, nests, iters;
  double r, T, tol;
  unsigned long long seed;
  const double* lo;      // [np]
  const double* hi;      // [np]
  const double* u;       // [m]
  const double* phiHat;  // [m][2]
  double* params_out;    // [restarts][np]
  double* fnval_out;     // [restarts]
  int* iters_out;        // [restarts] iterations actually run
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct Rng {  // one independent stream per (seed, restart, iteration, phase, nest, coordinate)
  unsigned long long key, ctr;
  __device__ Rng(unsigned long long seed, int restart, int it, int phase, int nest, int dim) {
    key = splitmix64(seed ^ splitmix64(((unsigned long long)restart << 32) | (unsigned)it));
    key = splitmix64(key ^ (((unsigned long long)phase << 48) | ((unsigned long long)nest << 16) | (unsigned)dim));
    ctr = 0;
  }
  __device__ double uniform() {  // (0, 1)
    const unsigned long long z = splitmix64(key + (ctr++) * 0xD1342543DE82EF95ull);
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double normal() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
};

__global__ void __launch_bounds__(32 * CK_WARPS) cuckoo_cf_kernel(const CuckooArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double pos[CK_MAX_NESTS][CK_MAX_NP], cand[CK_MAX_NESTS][CK_MAX_NP];
  __shared__ double val[CK_MAX_NESTS], cval[CK_MAX_NESTS];
  __shared__ double lo[CK_MAX_NP], hi[CK_MAX_NP];
  __shared__ int s_best, s_stop;
  ModelConsts* consts = reinterpret_cast<ModelConsts*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* terms = reinterpret_cast<double*>(smem_raw + sizeof(ModelConsts) * CK_WARPS) + (size_t)warp * a.m;
  const int restart = blockIdx.x, np = a.np, n = a.nests;
  if (tid < np) { lo[tid] = a.lo[tid]; hi[tid] = a.hi[tid]; }
  if (tid == 0) s_stop = 0;
  __syncthreads();

  auto eval = [&](const double* theta) {
    return warp_objfn_cf(a.base, a.tc, theta, a.m, a.u, a.phiHat, a.r, a.T, consts[warp], terms);
  };
  auto pick_best = [&]() {  // thread 0; ties -> lowest index
    int b = 0;
    for (int i = 1; i < n; ++i)
      if (val[i] < val[b]) b = i;
    s_best = b;
    if (val[b] < a.tol) s_stop = 1;
  };

  // initial nests: uniform in the box
  for (int i = warp; i < n; i += CK_WARPS) {
    if (lane < np) {
      Rng g(a.seed, restart, 0, 0, i, lane);
      pos[i][lane] = lo[lane] + (hi[lane] - lo[lane]) * g.uniform();
    }
    __syncwarp();
    const double v = eval(pos[i]);
    if (lane == 0) val[i] = v;
  }
  __syncthreads();
  if (tid == 0) pick_best();
  __syncthreads();

  const double sigma_u = 0.6965745025576968;  // Mantegna, beta = 1.5
  int it = 0;
  for (; it < a.iters && !s_stop; ++it) {
    // ---- phase 1: Levy flights relative to the best nest
    const int best = s_best;
    for (int i = warp; i < n; i += CK_WARPS) {
      if (lane < np) {
        Rng g(a.seed, restart, it + 1, 1, i, lane);
        const double un = g.normal() * sigma_u, vn = g.normal();
        const double step = un / pow(fabs(vn), 1.0 / 1.5);
        double x = pos[i][lane];
        x += 0.01 * step * (x - pos[best][lane]) * g.normal();
        cand[i][lane] = fmin(fmax(x, lo[lane]), hi[lane]);
      }
      __syncwarp();
      const double v = eval(cand[i]);
      if (lane == 0) cval[i] = v;
    }
    __syncthreads();
    for (int i = warp; i < n; i += CK_WARPS) {
      const bool better = cval[i] < val[i];  // NaN never accepted
      __syncwarp();
      if (better) {
        if (lane < np) pos[i][lane] = cand[i][lane];
        if (lane == 0) val[i] = cval[i];
      }
    }
    __syncthreads();
    // ---- phase 2: abandon a fraction pa of the coordinates (walk between two other nests)
    for (int i = warp; i < n; i += CK_WARPS) {
      Rng gi(a.seed, restart, it + 1, 2, i, CK_MAX_NP);
      const int j1 = (int)(gi.uniform() * n) % n, j2 = (int)(gi.uniform() * n) % n;
      const double scale = gi.uniform();
      if (lane < np) {
        Rng g(a.seed, restart, it + 1, 2, i, lane);
        double x = pos[i][lane];
        if (g.uniform() < 0.25) x += scale * (pos[j1][lane] - pos[j2][lane]);
        cand[i][lane] = fmin(fmax(x, lo[lane]), hi[lane]);
      }
      __syncwarp();
      const double v = eval(cand[i]);
      if (lane == 0) cval[i] = v;
    }
    __syncthreads();
    for (int i = warp; i < n; i += CK_WARPS) {
      const bool better = cval[i] < val[i];
      __syncwarp();
      if (better) {
        if (lane < np) pos[i][lane] = cand[i][lane];
        if (lane == 0) val[i] = cval[i];
      }
    }
    __syncthreads();
    if (tid == 0) pick_best();
    __syncthreads();
  }
  if (tid < np) a.params_out[(size_t)restart * np + tid] = pos[s_best][tid];
  if (tid == 0) {
    a.fnval_out[restart] = val[s_best];
    a.iters_out[restart] = it;
  }
}

}  // namespace

extern "C" int fop_calibrate_cf(fop_handle_t h, fop_model_t model, const double* lower,
                                const double* upper, const double* u, const double* phiHat, int m,
                                double r, double T, int nests, int iterations, double tol,
                                unsigned long long seed, int restarts, double* params_out,
                                double* fnval_out, int* iters_out, int* best_restart_out) {
  if (!h) return FOP_INVALID_ARG;
  const int np = num_params(model.base, model.time_change);
  if (np == 0) return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if (!lower || !upper || !u || !phiHat || !params_out || !fnval_out)
    return fail(h, FOP_INVALID_ARG, "null pointer");
  if (m < 1 || m > 2048 || nests < 2 || nests > CK_MAX_NESTS || iterations < 0 || restarts < 1)
    return fail(h, FOP_INVALID_ARG,
                "fop_calibrate_cf: need 1 <= m <= 2048, 2 <= nests <= %d, iterations >= 0, restarts >= 1",
                CK_MAX_NESTS);
  FOP_CUDA(h, cudaSetDevice(h->device));
  // the bounds may be host or device pointers like every other array argument: validate a host copy
  double lo_h[16], hi_h[16];
  FOP_CUDA(h, cudaMemcpy(lo_h, lower, (size_t)np * 8, cudaMemcpyDefault));
  FOP_CUDA(h, cudaMemcpy(hi_h, upper, (size_t)np * 8, cudaMemcpyDefault));
  for (int i = 0; i < np; ++i)
    if (!(hi_h[i] >= lo_h[i])) return fail(h, FOP_INVALID_ARG, "bound %d: upper < lower", i);
  Arena ar(h);
  FOP_TRY(ar.reserve(2 * Arena::pad((size_t)np * 8) + stage_bytes(u, (size_t)m * 8) +
                     stage_bytes(phiHat, (size_t)m * 16) + Arena::pad((size_t)restarts * np * 8) +
                     2 * Arena::pad((size_t)restarts * 8)));
  CuckooArgs a;
  a.base = model.base;  a.tc = model.time_change;  a.np = np;  a.m = m;  a.nests = nests;
  a.iters = iterations;  a.r = r;  a.T = T;  a.tol = tol;  a.seed = seed;
  double* d_lo = ar.alloc<double>(np);
  double* d_hi = ar.alloc<double>(np);
  FOP_CUDA(h, cudaMemcpyAsync(d_lo, lower, (size_t)np * 8, cudaMemcpyDefault, h->stream));
  FOP_CUDA(h, cudaMemcpyAsync(d_hi, upper, (size_t)np * 8, cudaMemcpyDefault, h->stream));
  a.lo = d_lo;  a.hi = d_hi;
  FOP_TRY(stage_in(h, ar, u, (size_t)m, &a.u));
  FOP_TRY(stage_in(h, ar, phiHat, (size_t)2 * m, &a.phiHat));
  a.params_out = ar.alloc<double>((size_t)restarts * np);
  a.fnval_out = ar.alloc<double>(restarts);
  a.iters_out = ar.alloc<int>(restarts);
  const size_t smem = sizeof(ModelConsts) * CK_WARPS + (size_t)CK_WARPS * m * 8;
  FOP_CUDA(h, cudaFuncSetAttribute(cuckoo_cf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope ps(h, "cuckoo_cf_kernel");
    cuckoo_cf_kernel<<<restarts, 32 * CK_WARPS, smem, h->stream>>>(a);
  }
  FOP_LAUNCH_CHECK(h);
  std::vector<double> fv(restarts);
  FOP_CUDA(h, cudaMemcpyAsync(fv.data(), a.fnval_out, (size_t)restarts * 8, cudaMemcpyDeviceToHost, h->stream));
  FOP_CUDA(h, cudaMemcpyAsync(params_out, a.params_out, (size_t)restarts * np * 8, cudaMemcpyDefault, h->stream));
  if (iters_out)
    FOP_CUDA(h, cudaMemcpyAsync(iters_out, a.iters_out, (size_t)restarts * 4, cudaMemcpyDefault, h->stream));
  FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  FOP_CUDA(h, cudaMemcpy(fnval_out, fv.data(), (size_t)restarts * 8, cudaMemcpyDefault));
  if (best_restart_out) {
    int b = 0;
    for (int i = 1; i < restarts; ++i)
      if (fv[i] < fv[b] || std::isnan(fv[b])) b = i;
    *best_restart_out = b;
  }
  return FOP_OK;
}
