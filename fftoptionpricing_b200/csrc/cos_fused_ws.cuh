// Fang-Oosterlee COS pricer, fused warp-specialised form (few strikes, one output kind).
//
// One persistent CTA per SM, 16 warps:
//   warps 0..11  producers: characteristic function -> COS coefficients, written straight into a
//                shared-memory ring of A stages (64 rows x 32 kk); one lane per (set, u_k) pair,
//                all maturities of the pair in registers (cf_prepare once, coeff_maturities).
//   warps 12..15 consumers: val = A . basis on the fp64 tensor cores (mma.sync m8n8k4), 16 rows x
//                (8 NB) strikes per warp, then the put/call-parity / Greek epilogue and the fused
//                squared-error row loss, stored from the accumulator registers.
// The basis stage (32 kk x strikes, shared by all rows) arrives by one cp.async.bulk per stage,
// completing on the same "full" mbarrier the producer lanes arrive on; consumers hand the slot
// back through an "empty" mbarrier.  Nothing but parameter sets in and prices / losses out
// touches HBM: the coefficient matrix of the split path (4 KiB per row, written and re-read)
// never exists.
//
// Why: both halves are bound by the same fp64 pipe (DFMA and DMMA share it: the mixed probe of
// fop_microbench_fp64 gives 35.3 TFLOP/s against 34.5 / 37.1 alone), and run back to back they
// leave it 35 % (CF: dependent exp/log/sincos chains, 16 warps) and 17 % (contraction) idle
// (profiles/r1_cos_c5_ncu.json).  Side by side on one SM the tensor warps fill the bubbles of
// the CF warps.
//
// Tile = SPT parameter sets x M maturities <= 64 rows (SPT = 64 / M); a stage has SPT x 16 (set,
// u) pairs = ceil(SPT / 2) warp items, dealt round-robin to the producer warps, which therefore
// work up to three stages ahead of the consumers (ring depth 5-6).
#pragma once
#include "cos_split_kernels.cuh"

namespace fop {

constexpr int FW_ROWS = 64;              // rows per tile (4 consumer warps x 16)
constexpr int FW_KC = 32;                // kk per stage
constexpr int FW_AS = FW_KC + 4;         // A row stride, doubles (conflict free, see DM_AS)
constexpr int FW_NPW = 12;               // producer warps
constexpr int FW_NCW = 4;                // consumer warps
constexpr int FW_THREADS = 32 * (FW_NPW + FW_NCW);
constexpr int FW_A_BYTES = FW_ROWS * FW_AS * 8;
constexpr int FW_MAX_M = 64;

__host__ __device__ constexpr int fw_b_bytes(int NB) { return FW_KC * dm_bs(NB) * 8; }
__host__ __device__ constexpr int fw_stage_bytes(int NB) { return FW_A_BYTES + fw_b_bytes(NB); }
__host__ __device__ constexpr int fw_tail_bytes(int NB) {  // barriers + disc[M] + strikes
  return 2 * 8 * 8 + FW_MAX_M * 8 + 8 * NB * 8;
}
__host__ inline int fw_depth(int NB, size_t smem_optin) {
  const int d = (int)((smem_optin - fw_tail_bytes(NB) - 128) / fw_stage_bytes(NB));
  return d > 8 ? 8 : d;
}

struct CosFusedArgs {
  CosCoeffArgs ca;   // base, tc, greek_mask (one bit), P, M, K, np, r, params, T, uk, vk (C unused)
  int kind, J, nchunks, SPT, depth;
  int vec2;          // prices may be stored as 16-byte pairs (J even, out 16-byte aligned)
  double S0;
  const ModelConsts* consts;  // [P]
  const double* basis;        // [nchunks * 32][dm_bs(NB)]
  const double* strikes;      // [J]
  double* out;                // [P*M][J] or nullptr
  const double* mkt;          // [M][J] or nullptr
  const double* w;            // [M][J] or nullptr
  double* rowloss;            // [P*M] or nullptr
};

static __global__ void cos_consts_kernel(int base, int tc, const double* __restrict__ params, int np, int P,
                                  ModelConsts* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) make_consts(base, tc, params + (size_t)p * np, out[p]);
}

// ---- mbarrier / bulk-copy primitives (PTX; CTA-local) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

template <int NB, int ILP>
__global__ void __launch_bounds__(FW_THREADS, 1) cos_fused_ws_kernel(const CosFusedArgs a) {
  constexpr int JB = 8 * NB, BS = dm_bs(NB);
  constexpr int STAGE = fw_stage_bytes(NB), BBYTES = fw_b_bytes(NB);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int D = a.depth;
  unsigned char* tail = smem_raw + (size_t)D * STAGE;
  const uint32_t bar0 = smem_u32(tail);             // full[d] at bar0 + 8 d, empty[d] at bar0 + 64 + 8 d
  double* s_disc = reinterpret_cast<double*>(tail + 128);
  double* s_strike = s_disc + FW_MAX_M;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = a.ca.M, SPT = a.SPT, nchunks = a.nchunks;
  const int nwi = (SPT + 1) >> 1;                    // warp items per stage
  const long long ntiles = ((long long)a.ca.P + SPT - 1) / SPT;
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const CosEpi epi = cos_epi_consts(a.kind, a.S0, a.ca.r);

  if (tid == 0) {
    for (int d = 0; d < D; ++d) {
      mbar_init(bar0 + 8 * d, nwi * 32 + 1);         // producer lanes + the expect_tx arrival
      mbar_init(bar0 + 64 + 8 * d, FW_NCW * 32);     // consumer lanes
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < M) s_disc[tid] = exp(-a.ca.r * __ldg(a.ca.T + tid)) * epi.inv;
  if (tid >= 64 && tid < 64 + JB) {
    const int j = tid - 64;
    s_strike[j] = j < a.J ? __ldg(a.strikes + j) : 0.0;
  }
  // rows no producer ever writes (SPT * M < 64, ragged last tile): defined values, not stale smem
  for (int d = 0; d < D; ++d)
    for (int i = tid; i < FW_A_BYTES / 16; i += FW_THREADS)
      reinterpret_cast<double2*>(smem_raw + (size_t)d * STAGE)[i] = make_double2(0.0, 0.0);
  __syncthreads();

  if (warp < FW_NPW) {
    // ------------------------------------------------------------------ producers
    const long long total = my_tiles * nchunks * nwi;
    const int ls_in = lane >> 4, ul = lane & 15;
    for (long long q = warp; q < total; q += FW_NPW) {
      const long long st = q / nwi;
      const int wi = (int)(q - st * nwi);
      const long long tl = st / nchunks;
      const int ch = (int)(st - tl * nchunks);
      const long long tile = blockIdx.x + tl * gridDim.x;
      const int slot = (int)(st % D);
      const uint32_t par = (uint32_t)((st / D) & 1);
      const uint32_t full = bar0 + 8 * slot;
      mbar_wait(bar0 + 64 + 8 * slot, par ^ 1u);
      unsigned char* stage = smem_raw + (size_t)slot * STAGE;
      if (wi == 0 && lane == 0) {
        mbar_arrive_expect_tx(full, BBYTES);
        bulk_g2s(smem_u32(stage + FW_A_BYTES), a.basis + (size_t)ch * FW_KC * BS, BBYTES, full);
      }
      const int ls = 2 * wi + ls_in;
      const long long set = tile * SPT + ls;
      if (ls < SPT && set < a.ca.P) {
        const int k = ch * (FW_KC / 2) + ul;
        double2* crow = reinterpret_cast<double2*>(stage) + (size_t)(ls * M) * (FW_AS / 2) + ul;
        if (k >= a.ca.K) {
          for (int m = 0; m < M; ++m) crow[(size_t)m * (FW_AS / 2)] = make_double2(0.0, 0.0);
        } else {
          const ModelConsts& mc = a.consts[set];
          const cd u = mk(0.0, __ldg(a.ca.uk + k));
          const double v = __ldg(a.ca.vk + k);
          CfPre pre;
          cf_prepare(a.ca.base, a.ca.tc, mc, u, a.ca.r, pre);
          const int mode = cf_mode(a.ca.tc, mc);
          if (mode == CF_MODE_CIR) coeff_maturities<CF_MODE_CIR, ILP>(a.ca, mc, pre, u, v, crow, FW_AS / 2);
          else if (mode == CF_MODE_LEVY) coeff_maturities<CF_MODE_LEVY, ILP>(a.ca, mc, pre, u, v, crow, FW_AS / 2);
          else coeff_maturities<CF_MODE_DET, 1>(a.ca, mc, pre, u, v, crow, FW_AS / 2);
        }
      }
      mbar_arrive(full);
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int cw = warp - FW_NPW;
    const int g = lane >> 2, c = lane & 3;
    const int rows_used = SPT * M;
    const bool vec2 = a.vec2 != 0;
    for (long long tl = 0; tl < my_tiles; ++tl) {
      const long long tile = blockIdx.x + tl * gridDim.x;
      double acc[2][NB][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int ch = 0; ch < nchunks; ++ch) {
        const long long st = tl * nchunks + ch;
        const int slot = (int)(st % D);
        const uint32_t par = (uint32_t)((st / D) & 1);
        mbar_wait(bar0 + 8 * slot, par);
        const unsigned char* stage = smem_raw + (size_t)slot * STAGE;
        const double* As = reinterpret_cast<const double*>(stage) + (cw * 16 + g) * FW_AS + c;
        const double* Bs = reinterpret_cast<const double*>(stage + FW_A_BYTES) + c * BS + g;
#pragma unroll
        for (int ks = 0; ks < FW_KC; ks += 4) {
          double af[2], bf[NB];
          af[0] = As[ks];
          af[1] = As[8 * FW_AS + ks];
#pragma unroll
          for (int j = 0; j < NB; ++j) bf[j] = Bs[ks * BS + 8 * j];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        mbar_arrive(bar0 + 64 + 8 * slot);
      }
      // epilogue from the accumulator fragment: lane holds rows g, g + 8 and columns 8 j + 2 c, + 1
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int rl = cw * 16 + 8 * i + g;
        const long long set = tile * SPT + rl / M;
        const bool valid = rl < rows_used && set < a.ca.P;
        const int m = rl % M;
        const long long row = tile * rows_used + rl;
        const double rs = s_disc[m];
        double* orow = a.out && valid ? a.out + (size_t)row * a.J : nullptr;
        const double* mrow = a.mkt && valid ? a.mkt + (size_t)m * a.J : nullptr;
        const double* wrow = a.w && valid ? a.w + (size_t)m * a.J : nullptr;
        double lsum = 0.0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int j0 = 8 * j + 2 * c;
          const double o0 = fma(acc[i][j][0] - epi.sub, rs * s_strike[j0], epi.add);
          const double o1 = fma(acc[i][j][1] - epi.sub, rs * s_strike[j0 + 1], epi.add);
          if (orow) {
            if (vec2) {
              if (j0 < a.J) *reinterpret_cast<double2*>(orow + j0) = make_double2(o0, o1);
            } else {
              if (j0 < a.J) orow[j0] = o0;
              if (j0 + 1 < a.J) orow[j0 + 1] = o1;
            }
          }
          if (mrow) {
            if (j0 < a.J) {
              const double d = o0 - __ldg(mrow + j0);
              lsum = fma((wrow ? __ldg(wrow + j0) : 1.0) * d, d, lsum);
            }
            if (j0 + 1 < a.J) {
              const double d = o1 - __ldg(mrow + j0 + 1);
              lsum = fma((wrow ? __ldg(wrow + j0 + 1) : 1.0) * d, d, lsum);
            }
          }
        }
        if (a.rowloss) {  // fixed order: columns within the lane, then lanes c = 0..3
          lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
          lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
          if (c == 0 && valid) a.rowloss[row] = lsum;
        }
      }
    }
  }
}

}  // namespace fop
