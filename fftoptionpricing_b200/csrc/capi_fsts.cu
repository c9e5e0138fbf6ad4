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
#include <cstdlib>

#include "cf_models.cuh"
#include "fft_engine.cuh"
#include "handle.cuh"

namespace {
using namespace fop;

struct FstsArgs {
  int base, tc, np, n, P, steps, american, payoff, kind;
  int tpb;  // paired kernel: slots per CTA
  double xmax, strike, r, Tcf, disc;  // Tcf = maturity fed to the CF (T or T/steps)
  const double* params;
  double* out;
};

__host__ __device__ inline size_t fsts_slot_bytes(int n) {
  const int N = 1 << n;
  return (size_t)fft_padded_size(N) * sizeof(cd) + (size_t)N * 8;
}

// Register-resident time loop.  Thread b of a set owns, at every step boundary, the 16 real values
// at positions b + S0 t (S0 = N/16) -- exactly the operands of the first forward pass and the
// results of the last inverse pass, so nothing is stored between steps.  Per step:
//   fwd pass 0 (regs -> smem) | middle passes (smem <-> regs) | last pass + multiply + first
//   inverse pass fused in registers | inverse middle passes | inverse pass 0 (smem -> regs) + max.
// Twiddles: one base W_L^q per radix-16 pass kept in registers, its powers by a depth-4 tree.
template <int LOGN>
__global__ void __launch_bounds__(256) fsts_kernel(const FstsArgs a, const double2* __restrict__ tw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // N is a template parameter: every stride, padded offset and pass count below is a compile-time
  // constant, so the time loop contains no integer division and addresses are base + immediate
  constexpr int n = LOGN, N = 1 << n, T = N >> 4;
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

  // plan: np16 radix-16 passes with strides S_p = N >> 4(p+1); when the last radix is 16 the
  // final one (S = 1) is the fused register pass, otherwise a small radix-rl pass is.
  constexpr int rl = (n & 3) == 0 ? 16 : (1 << (n & 3));
  constexpr int np16 = ((n + 3) >> 2) - (rl == 16 ? 0 : 1);
  constexpr int nsm = rl == 16 ? np16 - 1 : np16;  // radix-16 passes that go through shared memory
  constexpr int S0 = N >> 4;
  // twiddle bases (forward sign) of the shared-memory passes: w[p] = W_{16 S_p}^{b mod S_p}
  cd w[3];
  int base[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    constexpr int SS[3] = {N >> 4 > 0 ? N >> 4 : 1, N >> 8 > 0 ? N >> 8 : 1, N >> 12 > 0 ? N >> 12 : 1};
    const int S = SS[p];
    w[p] = p < nsm ? tw_load<-1>(tw, (b % S) * (N / (S << 4))) : mk(1.0, 0.0);
    base[p] = fft_base16(b, S);
  }

  for (int p0 = blockIdx.x * tpb; p0 < a.P; p0 += gridDim.x * tpb) {
    const int cnt = min(tpb, a.P - p0);
    const bool live = slot < cnt;
    __syncthreads();
    if ((int)threadIdx.x < cnt)
      make_consts(a.base, a.tc, a.params + (size_t)(p0 + threadIdx.x) * a.np, consts[threadIdx.x]);
    __syncthreads();

    // multiplier m at digit-reversed positions: one CF per loop trip into the (still unused)
    // value grid, then each thread pulls its 16 positions 16 b .. 16 b + 15 into registers
    {
      const ModelConsts& mc = consts[slot];
#pragma unroll 1
      for (int pos = b; pos < N; pos += T) {
        const int k = fft_inv_pos(pos, n);
        double wk = du * k;
        if (wk > vmax) wk -= 2.0 * vmax;  // getUDomain, OptionPricing.h:19-23 (strict >)
        const cd u = mk(0.0, wk);
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
      pay[j] = a.payoff == 1 ? (s < a.strike ? a.strike - s : 0.0) : (s > a.strike ? s - a.strike : 0.0);
    }
    __syncthreads();
    double vre[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) vre[t] = pay[b + S0 * t];

    // Idle slots of a tail block run the same instruction stream on their own (garbage) shared
    // memory so that every thread reaches every barrier; only global traffic is predicated.
    for (int st = 0; st < a.steps; ++st) {
      cd x[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) x[t] = mk(vre[t], 0.0);
      // ---- forward passes through shared memory
      if (nsm >= 1) {
        dif16_regs_pow<-1>(x, w[0]);
        store16_dft<-1>(sm, b, S0, x);
        __syncthreads();
#pragma unroll
        for (int p = 1; p < nsm; ++p) {
          const int S = N >> (4 * (p + 1));
          load16(sm, base[p], S, x);
          dif16_regs_pow<-1>(x, w[p]);
          store16_dft<-1>(sm, base[p], S, x);
          __syncthreads();
        }
        load16(sm, 16 * b, 1, x);
      }
      // ---- last forward pass, multiply, first inverse pass: all on the thread's 16 points
      if (rl == 16) {
        Dft<16, -1>::run(x);
        cd y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = x[Dft<16, -1>::out(k)] * mult[k];
        Dft<16, 1>::run(y);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = y[Dft<16, 1>::out(k)];
      } else {
        if (rl == 8) last_pass_regs<8, -1>(x);
        else if (rl == 4) last_pass_regs<4, -1>(x);
        else last_pass_regs<2, -1>(x);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = x[k] * mult[k];
        if (rl == 8) last_pass_regs<8, 1>(x);
        else if (rl == 4) last_pass_regs<4, 1>(x);
        else last_pass_regs<2, 1>(x);
      }
      // ---- inverse passes through shared memory (ascending stride), last one into registers
      if (nsm >= 1) {
        store16(sm, 16 * b, 1, x);
        __syncthreads();
#pragma unroll
        for (int p = nsm - 1; p >= 1; --p) {
          const int S = N >> (4 * (p + 1));
          load16(sm, base[p], S, x);
          dit16_regs_pow<1>(x, conj(w[p]));
          store16_dft<1>(sm, base[p], S, x);
          __syncthreads();
        }
        load16(sm, b, S0, x);
        dit16_regs_pow<1>(x, conj(w[0]));
#pragma unroll
        for (int k = 0; k < 16; ++k) vre[k] = x[Dft<16, 1>::out(k)].x;
        // no barrier: the next step's first pass rewrites exactly the positions this thread just read
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) vre[k] = x[k].x;
      }
      if (a.american) {
#pragma unroll
        for (int k = 0; k < 16; ++k) vre[k] = fmax(vre[k], pay[b + S0 * k]);
      }
    }

    if (live) {
      double* o = a.out + (size_t)(p0 + slot) * N;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int j = b + S0 * k;
        double v = vre[k];
        if (a.kind == 1 || a.kind == 2) {  // delta, gamma: divide by the asset (OptionPricing.h:226,275)
          const double s = a.strike * exp(-a.xmax + dx * j);
          v = a.kind == 1 ? v / s : v / (s * s);
        }
        o[j] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Paired variant: the value grids are REAL, so one complex transform carries two parameter sets,
// z = v_A + i v_B.  With Z = FFT(z), the spectra separate as V_A = (Z_k + conj Z_{N-k})/2,
// V_B = (Z_k - conj Z_{N-k})/(2i), and the step  v <- Re IFFT(V m)  for both sets at once is
//   W_k = A_k Z_k + B_k conj(Z_{N-k}),   A = (m_A + m_B)/2,  B = (m_A - m_B)/2,   z' = IFFT(W)
// (m Hermitian, which it is up to rounding because phi(-w) = conj phi(w); the two self-conjugate
// bins k = 0, N/2 get Im m = 0, which is what taking the real part does to them in the
// reference, OptionPricing.h:180,202).  Same passes, same registers (A where the unpaired kernel
// holds m; B is read from shared memory), one extra exchange of the 16 spectrum values of a
// thread with its partner positions pos(N - k) per step -- for twice the sets.
__host__ __device__ inline size_t fsts_pair_slot_bytes(int n) {
  const int N = 1 << n;
  return 2 * (size_t)fft_padded_size(N) * sizeof(cd) + (size_t)N * 8;
}

template <int LOGN>
__global__ void __launch_bounds__(256) fsts_pair_kernel(const FstsArgs a, const double2* __restrict__ tw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int n = LOGN, N = 1 << n, T = N >> 4;
  const int tpb = a.tpb;  // slots per CTA (<= fft_single_tpb(n)); a slot carries two parameter sets
  const int slot = threadIdx.x / T, b = threadIdx.x - slot * T;
  const size_t slot_bytes = fsts_pair_slot_bytes(n);
  cd* sm = reinterpret_cast<cd*>(smem_raw + (size_t)slot * slot_bytes);
  cd* bsm = sm + fft_padded_size(N);
  double* pay = reinterpret_cast<double*>(bsm + fft_padded_size(N));
  ModelConsts* consts = reinterpret_cast<ModelConsts*>(smem_raw + (size_t)tpb * slot_bytes);

  const double dx = (2.0 * a.xmax) / ((double)N - 1.0);
  const double vmax = M_PI / dx;
  const double du = 2.0 * vmax / N;
  const int greek = a.kind == 0 ? 0 : a.kind == 1 ? 1 : a.kind == 2 ? 2 : 3;
  const double scale = a.disc / (double)N;

  constexpr int rl = (n & 3) == 0 ? 16 : (1 << (n & 3));
  constexpr int np16 = ((n + 3) >> 2) - (rl == 16 ? 0 : 1);
  constexpr int nsm = rl == 16 ? np16 - 1 : np16;
  constexpr int S0 = N >> 4;
  cd w[3];
  int base[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    constexpr int SS[3] = {N >> 4 > 0 ? N >> 4 : 1, N >> 8 > 0 ? N >> 8 : 1, N >> 12 > 0 ? N >> 12 : 1};
    const int S = SS[p];
    w[p] = p < nsm ? tw_load<-1>(tw, (b % S) * (N / (S << 4))) : mk(1.0, 0.0);
    base[p] = fft_base16(b, S);
  }
  // padded shared-memory index of the partner pos(N - k) of each of the thread's 16 positions
  int pp[16];
#pragma unroll
  for (int k = 0; k < 16; ++k)
    pp[k] = fft_pad(fft_pos((N - fft_inv_pos(16 * b + k, n)) & (N - 1), n));

  for (int p0 = blockIdx.x * 2 * tpb; p0 < a.P; p0 += gridDim.x * 2 * tpb) {
    const int cnt = min(2 * tpb, a.P - p0);
    const int pA = p0 + 2 * slot, pB = pA + 1;
    const bool liveA = pA < a.P, liveB = pB < a.P;
    __syncthreads();
    if ((int)threadIdx.x < cnt)
      make_consts(a.base, a.tc, a.params + (size_t)(p0 + threadIdx.x) * a.np, consts[threadIdx.x]);
    __syncthreads();

    // multipliers of both sets at digit-reversed positions (an idle half repeats the last live set)
    {
      const int lastc = cnt - 1;
      const ModelConsts& mcA = consts[min(2 * slot, lastc)];
      const ModelConsts& mcB = consts[min(2 * slot + 1, lastc)];
#pragma unroll 1
      for (int pos = b; pos < N; pos += T) {
        const int k = fft_inv_pos(pos, n);
        double wk = du * k;
        if (wk > vmax) wk -= 2.0 * vmax;  // getUDomain, OptionPricing.h:19-23 (strict >)
        const cd u = mk(0.0, wk);
        const bool selfconj = k == 0 || 2 * k == N;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const ModelConsts& mc = h ? mcB : mcA;
          CfPre pre;
          cf_prepare(a.base, a.tc, mc, u, a.r, pre);
          const cd phi = cexp(cf_log_at(a.tc, mc, pre, a.Tcf));
          const cd m = option_transform(greek, phi, u, a.r);
          (h ? bsm : sm)[fft_pad(pos)] = mk(m.x * scale, selfconj ? 0.0 : m.y * scale);
        }
      }
    }
    __syncthreads();
    cd mult[16];  // A = (m_A + m_B)/2 in registers; B = (m_A - m_B)/2 back into bsm (own positions)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int idx = fft_pad_off(16 * b, k, 1);
      const cd m1 = sm[idx], m2 = bsm[idx];
      mult[k] = mk(0.5 * (m1.x + m2.x), 0.5 * (m1.y + m2.y));
      bsm[idx] = mk(0.5 * (m1.x - m2.x), 0.5 * (m1.y - m2.y));
    }
    __syncthreads();
    for (int j = b; j < N; j += T) {
      const double s = a.strike * exp(-a.xmax + dx * j);
      pay[j] = a.payoff == 1 ? (s < a.strike ? a.strike - s : 0.0) : (s > a.strike ? s - a.strike : 0.0);
    }
    __syncthreads();
    cd v[16];  // (v_A, v_B) at positions b + S0 t
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const double pj = pay[b + S0 * t];
      v[t] = mk(pj, pj);
    }

    for (int st = 0; st < a.steps; ++st) {
      cd x[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) x[t] = v[t];
      if (nsm >= 1) {
        dif16_regs_pow<-1>(x, w[0]);
        store16_dft<-1>(sm, b, S0, x);
        __syncthreads();
#pragma unroll
        for (int p = 1; p < nsm; ++p) {
          const int S = N >> (4 * (p + 1));
          load16(sm, base[p], S, x);
          dif16_regs_pow<-1>(x, w[p]);
          store16_dft<-1>(sm, base[p], S, x);
          __syncthreads();
        }
        load16(sm, 16 * b, 1, x);
      }
      // ---- last forward pass; exchange with the partner positions; combine; first inverse pass
      cd y[16];
      if (rl == 16) {
        Dft<16, -1>::run(x);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = x[Dft<16, -1>::out(k)];
      } else {
        if (rl == 8) last_pass_regs<8, -1>(x);
        else if (rl == 4) last_pass_regs<4, -1>(x);
        else last_pass_regs<2, -1>(x);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = x[k];
      }
      store16(sm, 16 * b, 1, y);  // Z at position 16 b + k (the thread's own positions: no hazard)
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const cd zp = sm[pp[k]];
        const cd bk = bsm[fft_pad_off(16 * b, k, 1)];
        // A Z + B conj(Zp)
        x[k] = mk(mult[k].x * y[k].x - mult[k].y * y[k].y + bk.x * zp.x + bk.y * zp.y,
                  mult[k].x * y[k].y + mult[k].y * y[k].x + bk.y * zp.x - bk.x * zp.y);
      }
      __syncthreads();  // partner reads done before the inverse passes overwrite the grid
      if (rl == 16) {
        Dft<16, 1>::run(x);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = x[Dft<16, 1>::out(k)];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = y[k];
      } else {
        if (rl == 8) last_pass_regs<8, 1>(x);
        else if (rl == 4) last_pass_regs<4, 1>(x);
        else last_pass_regs<2, 1>(x);
      }
      if (nsm >= 1) {
        store16(sm, 16 * b, 1, x);
        __syncthreads();
#pragma unroll
        for (int p = nsm - 1; p >= 1; --p) {
          const int S = N >> (4 * (p + 1));
          load16(sm, base[p], S, x);
          dit16_regs_pow<1>(x, conj(w[p]));
          store16_dft<1>(sm, base[p], S, x);
          __syncthreads();
        }
        load16(sm, b, S0, x);
        dit16_regs_pow<1>(x, conj(w[0]));
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = x[Dft<16, 1>::out(k)];
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = x[k];
      }
      if (a.american) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const double pj = pay[b + S0 * k];
          v[k] = mk(fmax(v[k].x, pj), fmax(v[k].y, pj));
        }
      }
    }

#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = b + S0 * k;
      double va = v[k].x, vb = v[k].y;
      if (a.kind == 1 || a.kind == 2) {  // delta, gamma: divide by the asset (OptionPricing.h:226,275)
        const double s = a.strike * exp(-a.xmax + dx * j);
        const double q = a.kind == 1 ? s : s * s;
        va = va / q;
        vb = vb / q;
      }
      if (liveA) a.out[(size_t)pA * N + j] = va;
      if (liveB) a.out[(size_t)pB * N + j] = vb;
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
// large grids (N > 4096) through the general FFT engine, capi_fft.cu
int fsts_large_run(fop_handle_t h, Arena& ar, int base, int tc, int np, const double* d_params, int P, int N,
                   double xmax, double strike, double r, double Tcf, double disc, int steps, int american,
                   int payoff, int kind, double* d_out);
size_t fsts_large_bytes(int N, int P);
}

extern "C" int fop_fsts(fop_handle_t h, fop_model_t model, const double* params, int P, int N,
                        double xmax, double strike, double r, double T, int steps, int american,
                        int payoff, int kind, double* out) {
  if (!h) return FOP_INVALID_ARG;
  const int np = num_params(model.base, model.time_change);
  if (np == 0) return fail(h, FOP_INVALID_ARG, "unknown model (%d,%d)", model.base, model.time_change);
  if (!params || !out) return fail(h, FOP_INVALID_ARG, "null pointer");
  if (P < 0 || N < 16 || (N & (N - 1)) || N > (1 << 20) || steps < 1 || kind < 0 || kind > 3 ||
      payoff < 0 || payoff > 1)
    return fail(h, FOP_INVALID_ARG, "fop_fsts: N power of two in [16, 2^20], steps >= 1");
  if ((steps > 1 || american) && (model.time_change != FOP_TC_NONE || kind != FOP_FSTS_PRICE))
    return fail(h, FOP_UNSUPPORTED,
                "fop_fsts: time stepping / early exercise needs a Levy model and kind = price");
  if (P == 0) return FOP_OK;
  FOP_CUDA(h, cudaSetDevice(h->device));
  const int n = ilog2(N);
  const bool out_host = !is_device_ptr(out);
  Arena ar(h);
  FOP_TRY(ar.reserve(stage_bytes(params, (size_t)P * np * 8) + (out_host ? Arena::pad((size_t)P * N * 8) : 0) +
                     (N > 4096 ? fsts_large_bytes(N, P) : 0)));
  const double* d_params = nullptr;
  FOP_TRY(stage_in(h, ar, params, (size_t)P * np, &d_params));
  double* d_out = out_host ? ar.alloc<double>((size_t)P * N) : out;
  const bool single = steps == 1 && !american;
  if (N > 4096) {  // beyond the register-resident kernels: general FFT engine
    const double Tcf = single ? T : T / steps;
    FOP_TRY(fsts_large_run(h, ar, model.base, model.time_change, np, d_params, P, N, xmax, strike, r, Tcf,
                           std::exp(-r * Tcf), steps, american ? 1 : 0, payoff, kind, d_out));
    if (out_host) {
      FOP_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)P * N * 8, cudaMemcpyDeviceToHost, h->stream));
      FOP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return FOP_OK;
  }
  const double2* tw = nullptr;
  FOP_TRY(get_twiddles_shared(h, N, &tw));

  FstsArgs a;
  a.base = model.base;  a.tc = model.time_change;  a.np = np;  a.n = n;  a.P = P;
  a.steps = steps;  a.american = american ? 1 : 0;  a.payoff = payoff;  a.kind = kind;
  a.xmax = xmax;  a.strike = strike;  a.r = r;
  a.Tcf = single ? T : T / steps;
  a.disc = std::exp(-r * a.Tcf);
  a.params = d_params;  a.out = d_out;
  int tpb = fft_single_tpb(n);
  // two parameter sets per complex transform whenever there are two (FOP_FSTS_PAIR=0: A/B runs)
  const char* pe = std::getenv("FOP_FSTS_PAIR");
  const bool pair = P >= 2 && !(pe && pe[0] == '0');
  if (pair)  // small grids: the per-set constants, not the grids, fill shared memory
    while (tpb > 1 && (size_t)tpb * (fsts_pair_slot_bytes(n) + 2 * sizeof(ModelConsts)) > (size_t)200 * 1024)
      tpb >>= 1;
  a.tpb = tpb;
  const size_t smem = pair ? (size_t)tpb * fsts_pair_slot_bytes(n) + (size_t)2 * tpb * sizeof(ModelConsts)
                           : (size_t)tpb * fsts_slot_bytes(n) + (size_t)tpb * sizeof(ModelConsts);
  void (*kern)(const FstsArgs, const double2*) = nullptr;
  switch (n) {
    case 4: kern = pair ? fsts_pair_kernel<4> : fsts_kernel<4>; break;
    case 5: kern = pair ? fsts_pair_kernel<5> : fsts_kernel<5>; break;
    case 6: kern = pair ? fsts_pair_kernel<6> : fsts_kernel<6>; break;
    case 7: kern = pair ? fsts_pair_kernel<7> : fsts_kernel<7>; break;
    case 8: kern = pair ? fsts_pair_kernel<8> : fsts_kernel<8>; break;
    case 9: kern = pair ? fsts_pair_kernel<9> : fsts_kernel<9>; break;
    case 10: kern = pair ? fsts_pair_kernel<10> : fsts_kernel<10>; break;
    case 11: kern = pair ? fsts_pair_kernel<11> : fsts_kernel<11>; break;
    default: kern = pair ? fsts_pair_kernel<12> : fsts_kernel<12>; break;
  }
  FOP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_cta = pair ? 2 * tpb : tpb;
  const int groups = (P + per_cta - 1) / per_cta;
  const int grid = std::min(groups, h->sm_count * 4);
  {
    ProfScope ps(h, "fsts_kernel");
    kern<<<grid, pair ? std::max(tpb * (N / 16), 1) : fft_single_threads(n), smem, h->stream>>>(a, tw);
  }
  FOP_LAUNCH_CHECK(h);
  if (out_host) {
    FOP_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)P * N * 8, cudaMemcpyDeviceToHost, h->stream));
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return FOP_OK;
}
