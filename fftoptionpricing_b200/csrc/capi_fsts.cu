// C ABI: fop_fsts -- Fourier space time stepping (OptionPricing.h:155-281)
//
// One group of N/16 threads per parameter set keeps the whole value grid in shared memory for
// all time steps; HBM is touched only for the parameters and the final N prices:
//   v <- payoff
//   repeat steps:  V = FFT(v)  (forward DIF, natural -> digit-reversed)
//                  V *= m      (m = enhCF(phi_dt(i w_k), i w_k) e^{-r dt}/N, held in REGISTERS in
//                               digit-reversed order: the same 16 values every step)
//                  v = Re IFFT(V)   (inverse DIT, digit-reversed -> natural)
//                  v = max(v, payoff)          (american)
// No reordering pass exists anywhere: DIF output order == DIT input order (fft_smem.cuh).
#include <algorithm>
#include <cmath>

#include "cf_models.cuh"
#include "fft_engine.cuh"
#include "handle.cuh"

namespace {
using namespace fop;

struct FstsArgs {
  int base, tc, np, n, P, steps, american, payoff, kind;
  double xmax, strike, r, Tcf, disc;  // Tcf = maturity fed to the CF (T or T/steps)
  const double* params;
  double* out;
};

__host__ __device__ inline size_t fsts_slot_bytes(int n) {
  const int N = 1 << n;
  return (size_t)fft_padded_size(N) * sizeof(cd) + (size_t)N * 8;
}

__global__ void __launch_bounds__(256) fsts_kernel(const FstsArgs a, const double2* __restrict__ tw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, N = 1 << n, T = N >> 4;
  const int tpb = fft_single_tpb(n);
  const int slot = threadIdx.x / T, b = threadIdx.x - slot * T;
  const size_t slot_bytes = fsts_slot_bytes(n);
  cd* sm = reinterpret_cast<cd*>(smem_raw + (size_t)slot * slot_bytes);
  double* pay = reinterpret_cast<double*>(smem_raw + (size_t)slot * slot_bytes +
                                          (size_t)fft_padded_size(N) * sizeof(cd));
  ModelConsts* consts = reinterpret_cast<ModelConsts*>(smem_raw + (size_t)tpb * slot_bytes);

  // grid of FSTSGeneric (OptionPricing.h:165-167)
  const double dx = (2.0 * a.xmax) / ((double)N - 1.0);
  const double vmax = M_PI / dx;
  const double du = 2.0 * vmax / N;
  const int greek = a.kind == 0 ? 0 : a.kind == 1 ? 1 : a.kind == 2 ? 2 : 3;
  const double scale = a.disc / (double)N;

  for (int p0 = blockIdx.x * tpb; p0 < a.P; p0 += gridDim.x * tpb) {
    const int cnt = min(tpb, a.P - p0);
    const bool live = slot < cnt;
    __syncthreads();
    if ((int)threadIdx.x < cnt)
      make_consts(a.base, a.tc, a.params + (size_t)(p0 + threadIdx.x) * a.np, consts[threadIdx.x]);
    __syncthreads();

    // multiplier m at digit-reversed positions: computed one CF per loop trip into the (still
    // unused) value grid, then each thread pulls its 16 positions 16 b .. 16 b + 15 into registers
    {
      const ModelConsts& mc = consts[slot];
#pragma unroll 1
      for (int pos = b; pos < N; pos += T) {
        const int k = fft_inv_pos(pos, n);
        double w = du * k;
        if (w > vmax) w -= 2.0 * vmax;  // getUDomain, OptionPricing.h:19-23 (strict >)
        const cd u = mk(0.0, w);
        CfPre pre;
        cf_prepare(a.base, a.tc, mc, u, a.r, pre);
        const cd phi = cexp(cf_log_at(a.tc, mc, pre, a.Tcf));
        const cd m = option_transform(greek, phi, u, a.r);
        sm[fft_pad(pos)] = mk(m.x * scale, m.y * scale);
      }
    }
    __syncthreads();
    cd mult[16];
    load16(sm, 16 * b, 1, mult);
    __syncthreads();
    // payoff on the log-asset grid x_j = -xmax + dx j, asset = strike e^{x_j}
    for (int j = b; j < N; j += T) {
      const double s = a.strike * exp(-a.xmax + dx * j);
      const double pv = a.payoff == 1 ? (s < a.strike ? a.strike - s : 0.0)
                                      : (s > a.strike ? s - a.strike : 0.0);
      pay[j] = pv;
      sm[fft_pad(j)] = mk(pv, 0.0);
    }
    __syncthreads();

    // Idle slots of a tail block run the same instruction stream on their own (garbage) shared
    // memory so that every thread reaches every barrier; only global traffic is predicated.
    const int np = fft_num_passes(n);
    for (int st = 0; st < a.steps; ++st) {
      for (int ps = 0; ps < np; ++ps) {
        fft_dif_pass<-1>(sm, b, n, ps, tw);
        __syncthreads();
      }
      {
        cd x[16];
        load16(sm, 16 * b, 1, x);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = x[i] * mult[i];
        store16(sm, 16 * b, 1, x);
      }
      __syncthreads();
      for (int ps = 0; ps < np; ++ps) {
        fft_dit_pass<1>(sm, b, n, ps, tw);
        __syncthreads();
      }
      for (int j = b; j < N; j += T) {
        double v = sm[fft_pad(j)].x;
        if (a.american) v = fmax(v, pay[j]);
        sm[fft_pad(j)] = mk(v, 0.0);
      }
      __syncthreads();
    }

    if (live) {
      double* o = a.out + (size_t)(p0 + slot) * N;
      for (int j = b; j < N; j += T) {
        double v = sm[fft_pad(j)].x;
        if (a.kind == 1 || a.kind == 2) {  // delta, gamma: divide by the asset (OptionPricing.h:226,275)
          const double s = a.strike * exp(-a.xmax + dx * j);
          v = a.kind == 1 ? v / s : v / (s * s);
        }
        o[j] = v;
      }
    }
  }
}

int ilog2(int N) {
  int n = 0;
  while ((1 << n) < N) ++n;
  return n;
}

}  // namespace

namespace fop {
int get_twiddles_shared(fop_handle_t h, int N, const double2** out);  // capi_fft.cu
}

extern "C" int fop_fsts(fop_handle_t h, fop_model_t model, const double* params, int P, int N,
                        double xmax, double strike, double r, double T, int steps, int american,
                        int payoff, int kind, double* out) {
  if (!h) return FOP_INVALID_ARG;
  const int np = num_params(model.base, model.time_change);
  if (np == 0) return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if (!params || !out) return fail(h, FOP_INVALID_ARG, "null pointer");
  if (P < 0 || N < 16 || (N & (N - 1)) || N > 4096 || steps < 1 || kind < 0 || kind > 3 ||
      payoff < 0 || payoff > 1)
    return fail(h, FOP_INVALID_ARG, "fop_fsts: N power of two in [16, 4096], steps >= 1");
  if ((steps > 1 || american) && (model.time_change != FOP_TC_NONE || kind != FOP_FSTS_PRICE))
    return fail(h, FOP_UNSUPPORTED,
                "fop_fsts: time stepping / early exercise needs a Levy model and kind = price");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const int n = ilog2(N);
  const bool out_host = !is_device_ptr(out);
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(params, (size_t)P * np * 8) + (out_host ? Arena::pad((size_t)P * N * 8) : 0)));
  const double* d_params = nullptr;
  FOP_TRY(stage_in(h, ar, params, (size_t)P * np, &d_params));
  double* d_out = out_host ? ar.alloc<double>((size_t)P * N) : out;
  const double2* tw = nullptr;
  FOP_TRY(get_twiddles_shared(h, N, &tw));

  const bool single = steps == 1 && !american;
  FstsArgs a;
  a.base = model.base;  a.tc = model.time_change;  a.np = np;  a.n = n;  a.P = P;
  a.steps = steps;  a.american = american ? 1 : 0;  a.payoff = payoff;  a.kind = kind;
  a.xmax = xmax;  a.strike = strike;  a.r = r;
  a.Tcf = single ? T : T / steps;
  a.disc = std::exp(-r * a.Tcf);
  a.params = d_params;  a.out = d_out;
  const int tpb = fft_single_tpb(n);
  const size_t smem = (size_t)tpb * fsts_slot_bytes(n) + (size_t)tpb * sizeof(ModelConsts);
  FOP_CUDA(h, cudaFuncSetAttribute(fsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = (P + tpb - 1) / tpb;
  const int grid = std::min(groups, h->sm_count * 4);
  {
    ProfScope ps(h, "fsts_kernel");
    fsts_kernel<<<grid, fft_single_threads(n), smem, h->stream>>>(a, tw);
  }
  FOP_LAUNCH_CHECK(h);
  if (out_host) {
    FOP_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)P * N * 8, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}
