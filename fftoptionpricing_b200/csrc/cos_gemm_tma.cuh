// COS contraction with Blackwell-native data movement (the default contraction kernel).
//
// Same tile and the same fp64 tensor-core inner product as cos_gemm_dmma_kernel
// (cos_split_kernels.cuh: 128 rows x (8 NB + EX) strikes per CTA, four warps of 32 rows,
// mma.sync.m8n8k4.f64, epilogue straight from the accumulator fragments), but the operands arrive by
// TMA instead of per-thread cp.async:
//   * A (coefficients C[rows][KK]) as 2-D tensor-map boxes of 128 rows x 16 kk with the 128-byte
//     swizzle -- the box is dense in shared memory and the swizzle makes the fragment reads
//     (8 rows x 4 consecutive doubles per quarter warp) conflict free, which the cp.async version
//     bought with 12.5 % padding; rows past the end of the matrix are zero-filled by the unit;
//   * B (basis[KK][Jp]) as boxes of 16 kk x (8 NB + 4) columns;
//   * a four-stage ring of 16-kk stages with one `full` mbarrier (TMA transaction bytes) and one
//     `empty` mbarrier (one arrival per warp) per stage.  One elected lane issues the loads three
//     stages ahead; no thread executes a cp.async, an address computation for it, or a CTA-wide
//     barrier inside the k loop (the cp.async kernel needed two __syncthreads per stage);
//   * persistent CTAs (two per SM): the ring runs across row tiles, so the loads of the next tile
//     are in flight while the epilogue of this one runs.
// tcgen05 / TMEM have no f64 kind, so the MMA itself stays mma.sync (DMMA).
#pragma once
#include <cuda.h>

#include "cos_split_kernels.cuh"

namespace fop {

constexpr int TM_KC = 16;   // kk per stage (= one 128-byte swizzle span of doubles)
constexpr int TM_ST = 4;    // stages
constexpr int TM_A_BYTES = DM_RT * TM_KC * 8;  // 16 KiB

__host__ __device__ constexpr int tm_b_bytes(int NB) { return ((TM_KC * dm_bs(NB) * 8 + 127) / 128) * 128; }
__host__ __device__ constexpr int tm_stage_bytes(int NB) { return TM_A_BYTES + ((tm_b_bytes(NB) + 1023) / 1024) * 1024; }
__host__ __device__ inline size_t cos_tma_smem(int NB) { return (size_t)TM_ST * tm_stage_bytes(NB) + 1024; }
// + market / weight tables of one strike block [2][M][8 NB + 2] when two CTAs still fit an SM
__host__ inline size_t cos_tma_tab_bytes(int NB, int M) { return (size_t)2 * M * (8 * NB + 2) * 8; }

__device__ __forceinline__ uint32_t tm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tm_smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tm_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tm_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TMWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TMDONE_%=;\n"
      "bra TMWAIT_%=;\n"
      "TMDONE_%=:\n"
      "}" ::"r"(tm_smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          tm_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(tm_smem_u32(bar))
      : "memory");
}

template <int NB, int EX = 0, int MB = 4>
__global__ void __launch_bounds__(32 * (DM_RT / (8 * MB)), 2)
    cos_gemm_tma_kernel(const CosGemmArgs a, const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB) {
  static_assert(EX >= 0 && EX <= 2, "at most two remainder columns fit the tile padding");
  constexpr int NT = 32 * (DM_RT / (8 * MB));
  constexpr int JB = 8 * NB, BS = dm_bs(NB);
  constexpr int STAGE = tm_stage_bytes(NB);
  constexpr uint32_t TX_BYTES = TM_A_BYTES + TM_KC * BS * 8;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_full[TM_ST], s_empty[TM_ST];
  __shared__ double s_rowscale[DM_RT];
  __shared__ int s_rowm[DM_RT];
  __shared__ double s_strike[JB];
  // market / weight tables of this strike block [M][JB + 2] when they fit (a.tab_smem; the host sizes
  // the dynamic shared memory): the loss epilogue's per-output look-ups were dependent L1/L2 loads
  // (16 % of the warp samples, profiles/r2_cos_gemm_tma_ncu.json)
  // 1024-byte aligned stage ring (the 128-byte swizzle pattern repeats every 1024 bytes)
  unsigned char* ring = smem_dyn + ((1024u - (tm_smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr int TW = JB + 2;  // table row pitch (doubles, even)
  double* s_mkt = reinterpret_cast<double*>(ring + (size_t)TM_ST * STAGE);
  double* s_w = s_mkt + (size_t)a.M * TW;
  const bool tab = a.tab_smem != 0 && a.mkt != nullptr;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int blk = blockIdx.y;
  const int nch = a.KK / TM_KC;
  const CosEpi epi = cos_epi_consts(a.kind, a.S0, a.r);
  const int jvalid = min(JB, a.J - blk * JB);
  // persistent: CTA x takes the row tiles x, x + gridDim.x, ...; the stage ring runs ACROSS tiles (one
  // global chunk counter), so the first stages of the next tile are already in flight while this
  // tile's epilogue runs
  const long long ntiles = (a.rows + DM_RT - 1) / DM_RT;
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_ch = my_tiles * nch;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TM_ST; ++s) {
      tm_mbar_init(&s_full[s], 1);
      tm_mbar_init(&s_empty[s], DM_RT / (8 * MB));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < JB) {
    const int j = blk * JB + tid;
    s_strike[tid] = j < a.J ? __ldg(a.strikes + j) : 0.0;
  }
  if (tab) {
    for (int i = tid; i < a.M * TW; i += NT) {
      const int m = i / TW, jj = i - m * TW, j = blk * JB + jj;
      const bool in = j < a.J && jj < JB + EX;
      s_mkt[i] = in ? __ldg(a.mkt + (size_t)m * a.J + j) : 0.0;
      s_w[i] = in ? (a.w ? __ldg(a.w + (size_t)m * a.J + j) : 1.0) : 0.0;
    }
  }
  __syncthreads();

  auto issue = [&](long long gch) {  // one elected lane: both boxes of global chunk gch into its ring slot
    const int s = (int)(gch % TM_ST);
    const long long t = gch / nch;
    const int ch = (int)(gch - t * nch);
    const long long r0 = ((long long)blockIdx.x + t * gridDim.x) * DM_RT;
    unsigned char* st = ring + (size_t)s * STAGE;
    tm_mbar_expect_tx(&s_full[s], TX_BYTES);
    tma_load_2d(st, &tmA, ch * TM_KC, (int)r0, &s_full[s]);
    tma_load_2d(st + TM_A_BYTES, &tmB, blk * JB, ch * TM_KC, &s_full[s]);
  };
  if (tid == 0) {
    for (long long gch = 0; gch < TM_ST - 1 && gch < total_ch; ++gch) issue(gch);
  }

  const int wrow = warp * (8 * MB);
  // swizzled byte offset of A[wrow + 8 i + g][ks + c] inside a stage:
  //   row * 128 + ((chunk ^ (row & 7)) << 4) + (col & 1) * 8,  chunk = col >> 1,  row & 7 = g
  int aoff[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) aoff[q] = (wrow + g) * 128 + ((((4 * q + c) >> 1) ^ g) << 4) + (c & 1) * 8;
  const int boff = (c * BS + g) * 8;

  long long gch = 0;
  for (long long t = 0; t < my_tiles; ++t) {
    const long long row0 = ((long long)blockIdx.x + t * gridDim.x) * DM_RT;
    const int nrows = (int)min((long long)DM_RT, a.rows - row0);
    __syncthreads();  // the previous tile's epilogue has read the row tables
    if (tid < nrows) {   // NT >= 128 = DM_RT
      const int m = (int)((row0 + tid) % a.M);
      s_rowm[tid] = m;
      s_rowscale[tid] = exp(-a.r * __ldg(a.T + m)) * epi.inv;
    }
    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double ex[MB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i) ex[i][0] = ex[i][1] = 0.0;

    for (int ch = 0; ch < nch; ++ch, ++gch) {
      const int s = (int)(gch % TM_ST);
      if (tid == 0) {  // keep three stages in flight (into the next tile at the end of this one)
        const long long nxt = gch + TM_ST - 1;
        if (nxt < total_ch) {
          if (nxt >= TM_ST) tm_mbar_wait(&s_empty[nxt % TM_ST], (uint32_t)(((nxt / TM_ST) - 1) & 1));
          issue(nxt);
        }
      }
      __syncwarp();
      tm_mbar_wait(&s_full[s], (uint32_t)((gch / TM_ST) & 1));
      const unsigned char* As = ring + (size_t)s * STAGE;
      const unsigned char* Bs = As + TM_A_BYTES + boff;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ks = 4 * q;
        double af[MB], bf[NB];
#pragma unroll
        for (int i = 0; i < MB; ++i) af[i] = *reinterpret_cast<const double*>(As + aoff[q] + i * 1024);
#pragma unroll
        for (int j = 0; j < NB; ++j) bf[j] = *reinterpret_cast<const double*>(Bs + (ks * BS + 8 * j) * 8);
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        if (EX > 0) {  // A[row][ks + c] * Bas[ks + c][JB + e]
          const double2 bx = *reinterpret_cast<const double2*>(Bs + ((ks * BS + JB) - g) * 8);
#pragma unroll
          for (int i = 0; i < MB; ++i) {
            ex[i][0] = fma(af[i], bx.x, ex[i][0]);
            if (EX > 1) ex[i][1] = fma(af[i], bx.y, ex[i][1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) tm_mbar_arrive(&s_empty[s]);
    }
    __syncthreads();  // row tables of this tile are complete (written before the k loop by other warps)

    // ---- epilogue: identical to cos_gemm_dmma_kernel (from the accumulator fragments)
    const bool vec2 = a.vec2 != 0, mvec2 = a.mvec2 != 0;
#pragma unroll
    for (int i = 0; i < MB; ++i) {
      const int rr = wrow + 8 * i + g;
      const bool valid = rr < nrows;
      const long long row = row0 + rr;
      const double rs = s_rowscale[valid ? rr : 0];
      const int m = s_rowm[valid ? rr : 0];
      double* orow = a.out && valid ? a.out + (size_t)row * a.J + (size_t)blk * JB : nullptr;
      const double* mrow = a.mkt && valid ? a.mkt + (size_t)m * a.J + (size_t)blk * JB : nullptr;
      const double* wrw = a.w && valid ? a.w + (size_t)m * a.J + (size_t)blk * JB : nullptr;
      double lsum = 0.0;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int j0 = 8 * j + 2 * c;
        const double o0 = fma(acc[i][j][0] - epi.sub, rs * s_strike[j0], epi.add);
        const double o1 = fma(acc[i][j][1] - epi.sub, rs * s_strike[j0 + 1], epi.add);
        if (orow) {
          if (vec2) {
            if (j0 < jvalid) *reinterpret_cast<double2*>(orow + j0) = make_double2(o0, o1);
          } else {
            if (j0 < jvalid) orow[j0] = o0;
            if (j0 + 1 < jvalid) orow[j0 + 1] = o1;
          }
        }
        if (tab) {  // zero weight outside the valid columns: no predicate needed
          if (valid) {
            const double2 mk2 = *reinterpret_cast<const double2*>(s_mkt + m * TW + j0);
            const double2 w2 = *reinterpret_cast<const double2*>(s_w + m * TW + j0);
            const double d0 = o0 - mk2.x, d1 = o1 - mk2.y;
            lsum = fma(w2.x * d0, d0, lsum);
            lsum = fma(w2.y * d1, d1, lsum);
          }
        } else if (mrow) {
          if (mvec2) {
            if (j0 < jvalid) {
              const double2 mk2 = __ldg(reinterpret_cast<const double2*>(mrow + j0));
              const double2 w2 = wrw ? __ldg(reinterpret_cast<const double2*>(wrw + j0)) : make_double2(1.0, 1.0);
              const double d0 = o0 - mk2.x, d1 = o1 - mk2.y;
              lsum = fma(w2.x * d0, d0, lsum);
              lsum = fma(w2.y * d1, d1, lsum);
            }
          } else {
            if (j0 < jvalid) {
              const double d = o0 - __ldg(mrow + j0);
              lsum = fma((wrw ? __ldg(wrw + j0) : 1.0) * d, d, lsum);
            }
            if (j0 + 1 < jvalid) {
              const double d = o1 - __ldg(mrow + j0 + 1);
              lsum = fma((wrw ? __ldg(wrw + j0 + 1) : 1.0) * d, d, lsum);
            }
          }
        }
      }
      if (EX > 0) {
        double e0 = ex[i][0], e1 = ex[i][1];
        e0 += __shfl_xor_sync(0xffffffffu, e0, 1);
        e0 += __shfl_xor_sync(0xffffffffu, e0, 2);
        if (EX > 1) {
          e1 += __shfl_xor_sync(0xffffffffu, e1, 1);
          e1 += __shfl_xor_sync(0xffffffffu, e1, 2);
        }
        if (c < EX) {
          const int jx = JB + c;
          const double o = fma((c == 0 ? e0 : e1) - epi.sub, rs * __ldg(a.strikes + jx), epi.add);
          if (orow) orow[jx] = o;
          if (tab) {
            if (valid) {
              const double d = o - s_mkt[m * TW + jx];
              lsum = fma(s_w[m * TW + jx] * d, d, lsum);
            }
          } else if (mrow) {
            const double d = o - __ldg(mrow + jx);
            lsum = fma((wrw ? __ldg(wrw + jx) : 1.0) * d, d, lsum);
          }
        }
      }
      if (a.rowloss) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
        if (c == 0 && valid) a.rowloss[(size_t)row * a.nblk + blk] = lsum;
      }
    }
  }
}

// table of the instantiated TMA contraction kernels (capi_cos_gemm_tma.cu)
struct GemmTmaVariant {
  int NB, EX;
  void (*kern)(const CosGemmArgs, const CUtensorMap, const CUtensorMap);
  void (*kern2)(const CosGemmArgs, const CUtensorMap, const CUtensorMap);  // 16-row warp tiles (8 warps)
  void (*kern1)(const CosGemmArgs, const CUtensorMap, const CUtensorMap);  // 8-row warp tiles (16 warps), or null
};
const GemmTmaVariant* gemm_tma_variants(int* count);

}  // namespace fop
