// Fang-Oosterlee COS pricer, two-kernel form (the default path).
//
//   1. cos_coeff_kernel   one thread per (parameter set, u_k) pair: characteristic function once
//                         per pair (cf_prepare) + once per maturity (cf_log_at), option transform,
//                         times V_k cp -> COS coefficients C[row][2k], C[row][2k+1]
//                         (row = set * M + maturity).  Register-light => high occupancy, which is
//                         what the long dependent fp64 chains of exp/log/sincos/atan2 need.
//   2. cos_gemm_dmma_kernel  val[row][j] = sum_kk C[row][kk] Bas[kk][j]: tall-skinny real fp64
//                         GEMM (A = C streamed from L2/HBM, B = basis shared by every row) on the
//                         fp64 tensor cores, 128 x (8 NB) CTA tiles, cp.async double buffering,
//                         then the put/call-parity / Greek epilogue and the optional fused
//                         squared-error row loss.
// C makes one round trip through L2/HBM (4 KiB per row at numU = 256): at the fp64 rate this is
// ~20 % of HBM bandwidth and fully hidden; in exchange each kernel runs at the occupancy / tile
// shape that suits it (the single fused kernel in cos_kernels.cuh sat at 33 % fp64-pipe
// utilisation with 8 warps/SM -- profiles/r1_cos_fused_v0_ncu_summary.json).
#pragma once
#include "cos_kernels.cuh"
#include "cos_coeff.cuh"

namespace fop {

// BASE / TC < 0: taken from the arguments at run time (generic kernel); >= 0: compile-time model
// (specialised kernels, capi_cos_coeff.cu)
template <int ILP, int NT = 256, int MINB = 2, int BASE = -1, int TC = -1>
__global__ void __launch_bounds__(NT, MINB) cos_coeff_kernel(const CosCoeffArgs a) {
  __shared__ ModelConsts consts[COEFF_MAX_SPB];
  const int tid = threadIdx.x;
  const int K = a.K, M = a.M;
  const int base = BASE >= 0 ? BASE : a.base, tc = TC >= 0 ? TC : a.tc;
  const int ngroups = (a.P + a.spb - 1) / a.spb;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int p0 = grp * a.spb;
    const int nsets = min(a.spb, a.P - p0);
    __syncthreads();
    if (tid < nsets) make_consts(base, tc, a.params + (size_t)(p0 + tid) * a.np, consts[tid]);
    __syncthreads();
    const int Kp = a.Kp;
    for (int idx = tid; idx < nsets * Kp; idx += NT) {
      const int s = idx / Kp, k = idx - s * Kp;
      double2* crow = reinterpret_cast<double2*>(a.C) + (size_t)(p0 + s) * M * Kp + k;
      if (k >= K) {  // zero padding so the contraction can run full stages
        double2* cg = crow;
        for (int g = 0; g < 4; ++g) {
          if (!(a.greek_mask >> g & 1)) continue;
          for (int m = 0; m < M; ++m) cg[(size_t)m * Kp] = make_double2(0.0, 0.0);
          cg += a.plane / 2;
        }
        continue;
      }
      const ModelConsts& mc = consts[s];
      const cd u = mk(0.0, __ldg(a.uk + k));
      const double v = __ldg(a.vk + k);
      CfPre pre;
      cf_prepare(base, tc, mc, u, a.r, pre);
      const int mode = cf_mode(tc, mc);
      if (mode == CF_MODE_CIR) coeff_maturities<CF_MODE_CIR, ILP>(a, mc, pre, u, v, crow, Kp);
      else if (mode == CF_MODE_LEVY) coeff_maturities<CF_MODE_LEVY, ILP>(a, mc, pre, u, v, crow, Kp);
      else coeff_maturities<CF_MODE_DET, 1>(a, mc, pre, u, v, crow, Kp);
    }
  }
}

// Coefficients from caller-sampled characteristic-function values (fop_cos_cf_samples: the
// reference's arbitrary CF callable, FangOostGeneric OptionPricing.h:325-333, evaluated by the host
// at i u_k): option transform, times V_k cp, zero padding -- then the same contraction.
struct CosSampleArgs {
  int greek_mask;
  size_t plane;
  long long rows;      // sets * maturities of this chunk
  int K, Kp;
  double r;
  const double* phi;   // [rows][K][2]
  const double* uk;
  const double* vk;
  double* C;
};
static __global__ void cos_coeff_samples_kernel(const CosSampleArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * a.Kp) return;
  const long long row = idx / a.Kp;
  const int k = (int)(idx - row * a.Kp);
  double2* cg = reinterpret_cast<double2*>(a.C) + row * a.Kp + k;
  const bool pad = k >= a.K;
  cd phi = mk(0.0, 0.0), u = mk(0.0, 0.0);
  double v = 0.0;
  if (!pad) {
    const double2 f = reinterpret_cast<const double2*>(a.phi)[row * a.K + k];
    phi = mk(f.x, f.y);
    u = mk(0.0, a.uk[k]);
    v = a.vk[k];
  }
  for (int g = 0; g < 4; ++g) {
    if (!(a.greek_mask >> g & 1)) continue;
    if (pad) {
      *cg = make_double2(0.0, 0.0);
    } else {
      const cd t = option_transform(g, phi, u, a.r);
      *cg = make_double2(t.x * v, t.y * v);
    }
    cg += a.plane / 2;
  }
}

struct CosGemmArgs {
  int kind;
  int M, J, KK;     // KK = 2 numU
  int Jp, nblk;     // padded strike count = nblk * JB
  long long rows;   // P * M
  double r, S0;
  const double* C;        // [rows][KK]
  const double* basis;    // [KK][Jp]
  const double* T;        // [M]
  const double* strikes;  // [J]
  double* out;            // [rows][J] or nullptr
  const double* mkt;      // [M][J] or nullptr
  const double* w;        // [M][J] or nullptr
  double* rowloss;        // [rows][nblk] or nullptr
  int vec2;               // prices may be stored as 16-byte pairs (J even, out 16-byte aligned)
  int mvec2;              // mkt / w may be loaded as 16-byte pairs (J even, tables 16-byte aligned)
  int tab_smem;           // TMA kernel: market / weight tables staged in shared memory
};

// ---------------------------------------------------------------------------------------------
// fp64 tensor-core contraction (mma.sync.m8n8k4.f64, "DMMA").
//
// Why: ncu on the first, DFMA register-tiled version of this kernel
// (profiles/r1_cos_c5_split_v1_ncu.json) showed it to be shared-memory bound -- every LDS.128 costs 4 wavefronts whether or not lanes broadcast, a
// 4 x 18 register tile needs 22 x 16 B per lane per 144 FMAs = 1.22x the 128 B/clk the LSU can
// deliver at the fp64 rate, and a tile large enough (8 x 18) does not fit in 255 registers.  The
// MMA unit shares the A / B fragments across lanes in hardware: a 32 x 72 warp tile needs
// 13 x 8 B per lane per 288 FMAs (0.36 B/FMA instead of 2.4), with the same 72 accumulators, and
// DMMA issues at the full 64 FMA/clk/SM (37.1 vs 34.5 TFLOP/s measured for DFMA).
//
// Fragment layout of mma.m8n8k4 (g = lane / 4, c = lane % 4):
//   A (8 x 4, row)  a     = A[g][c]
//   B (4 x 8, col)  b     = B[c][g]
//   C (8 x 8)       d0,d1 = C[g][2c], C[g][2c+1]
constexpr int DM_KC = 32;            // kk per stage
constexpr int DM_RT = 128;           // rows per CTA (4 warps x 32 rows)
constexpr int DM_AS = DM_KC + 4;     // A row stride (288 B: rows shift by 32 B -> conflict free)

__host__ __device__ constexpr int dm_bs(int NB) { return 8 * NB + 4; }  // B row stride (doubles)
__host__ __device__ inline size_t cos_dmma_smem(int NB) {
  const size_t stage = (size_t)DM_RT * DM_AS * 8 + (size_t)DM_KC * dm_bs(NB) * 8;
  return 2 * stage;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// C rows are padded to KKp = multiple of DM_KC (zero filled), basis rows likewise, so every
// stage is full and all loop bounds are compile-time constants.
// MB = m-blocks (of 8 rows) per warp: MB = 4 -> 4 warps x 32 rows (72 accumulators/lane at NB = 9),
// MB = 2 -> 8 warps x 16 rows (36 accumulators/lane, <= 128 registers -> 16 warps/SM instead of 8,
// which keeps the DMMA pipe busier across barriers / epilogues at 1.7x the smem traffic per MMA).
// EX = 1 or 2 remainder columns (J = 8 NB + EX, one strike block): they live in the padding of the
// basis tile and are contracted with plain DFMAs on the A fragments the lane already holds (the
// lane's share is kk = c mod 4; the four lanes of a row are summed after the main loop) -- 66
// strikes cost 8 n-blocks of DMMA plus 8 DFMAs per k-step instead of 9 n-blocks.
template <int NB, int MB, int EX = 0>
__global__ void __launch_bounds__(32 * (DM_RT / (8 * MB)), 2) cos_gemm_dmma_kernel(const CosGemmArgs a) {
  static_assert(EX >= 0 && EX <= 2, "at most two remainder columns fit the tile padding");
  constexpr int NT = 32 * (DM_RT / (8 * MB));
  constexpr int JB = 8 * NB, BS = dm_bs(NB);
  constexpr size_t A_BYTES = (size_t)DM_RT * DM_AS * 8;
  constexpr size_t STAGE = A_BYTES + (size_t)DM_KC * BS * 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_rowscale[DM_RT];  // disc(row) * inv
  __shared__ int s_rowm[DM_RT];         // maturity index of the row
  __shared__ double s_strike[JB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const long long row0 = (long long)blockIdx.x * DM_RT;
  const int blk = blockIdx.y;
  const int nrows = (int)min((long long)DM_RT, a.rows - row0);
  const int KKp = a.KK;  // padded
  const int nchunks = KKp / DM_KC;
  const CosEpi epi = cos_epi_consts(a.kind, a.S0, a.r);
  const int jvalid = min(JB, a.J - blk * JB);  // columns of the DMMA tile (the EX remainder is extra)

  // cp.async plan (fully coalesced: consecutive threads copy consecutive 16-byte pieces).
  //   A: piece q = tid + NT i  ->  row (tid >> 4) + (NT/16) i, piece tid & 15 of the row's 256 B
  //   B: piece q = tid + NT i  ->  basis row q / BPC, piece q % BPC   (offsets precomputed)
  constexpr int BPC = JB / 2 + (EX ? 1 : 0);           // 16-byte pieces per basis row
  constexpr int NBI = (DM_KC * BPC + NT - 1) / NT;     // B pieces per thread
  constexpr int ARS = NT / 16;                         // A rows covered per pass
  constexpr int API = DM_RT / ARS;                     // A passes
  const bool full = nrows == DM_RT;
  const double* abase = a.C + (size_t)(row0 + (tid >> 4)) * KKp + 2 * (tid & 15);
  double* adst = reinterpret_cast<double*>(smem_raw) + (tid >> 4) * DM_AS + 2 * (tid & 15);
  int bsoff[NBI], bgoff[NBI];
#pragma unroll
  for (int i = 0; i < NBI; ++i) {
    const int q = tid + NT * i;
    const int rr = q / BPC, cc = q - rr * BPC;
    bsoff[i] = q < DM_KC * BPC ? rr * BS + 2 * cc : -1;
    bgoff[i] = rr * a.Jp + 2 * cc;
  }
  const double* bbase = a.basis + (size_t)blk * JB;
  double* bdst = reinterpret_cast<double*>(smem_raw + A_BYTES);
  auto issue = [&](int ch, int buf) {
    const size_t so = (size_t)buf * (STAGE / 8);
    const int kk0 = ch * DM_KC;
    if (full) {
#pragma unroll
      for (int i = 0; i < API; ++i)
        cp_async16(adst + so + (ARS * i) * DM_AS, abase + kk0 + (size_t)(ARS * i) * KKp);
    } else {  // last row tile: clamp the source row (rows past the end are never stored)
#pragma unroll 1
      for (int i = 0; i < API; ++i) {
        const int rr = (tid >> 4) + ARS * i;
        const int rs = rr < nrows ? rr : nrows - 1;
        cp_async16(adst + so + (ARS * i) * DM_AS,
                   a.C + (size_t)(row0 + rs) * KKp + kk0 + 2 * (tid & 15));
      }
    }
    const double* bs = bbase + (size_t)kk0 * a.Jp;
#pragma unroll
    for (int i = 0; i < NBI; ++i)
      if (bsoff[i] >= 0) cp_async16(bdst + so + bsoff[i], bs + bgoff[i]);
    cp_async_commit();
  };
  issue(0, 0);

  // per-row / per-strike epilogue scalars (overlaps the first stage's latency)
  if (tid < nrows) {
    const int m = (int)((row0 + tid) % a.M);
    s_rowm[tid] = m;
    s_rowscale[tid] = exp(-a.r * __ldg(a.T + m)) * epi.inv;
  }
  if (tid < JB) {
    const int j = blk * JB + tid;
    s_strike[tid] = j < a.J ? __ldg(a.strikes + j) : 0.0;
  }

  double acc[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ex[MB][2];  // remainder columns JB, JB + 1: this lane's share of the kk sum
#pragma unroll
  for (int i = 0; i < MB; ++i) ex[i][0] = ex[i][1] = 0.0;

  const int wrow = warp * (8 * MB);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) {
      issue(ch + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* As = reinterpret_cast<const double*>(smem_raw + (size_t)buf * STAGE) + (wrow + g) * DM_AS + c;
    const double* Bs = reinterpret_cast<const double*>(smem_raw + (size_t)buf * STAGE + A_BYTES) + c * BS + g;
#pragma unroll
    for (int ks = 0; ks < DM_KC; ks += 4) {
      double af[MB], bf[NB];
#pragma unroll
      for (int i = 0; i < MB; ++i) af[i] = As[(8 * i) * DM_AS + ks];
#pragma unroll
      for (int j = 0; j < NB; ++j) bf[j] = Bs[ks * BS + 8 * j];
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      if (EX > 0) {  // A[row][ks + c] * Bas[ks + c][JB + e]
        const double2 bx = *reinterpret_cast<const double2*>(Bs - g + ks * BS + JB);
#pragma unroll
        for (int i = 0; i < MB; ++i) {
          ex[i][0] = fma(af[i], bx.x, ex[i][0]);
          if (EX > 1) ex[i][1] = fma(af[i], bx.y, ex[i][1]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue straight from the accumulator fragment: the lane holds rows wrow + 8 i + g and
  // columns 8 j + 2 c, + 1.  No shared-memory staging and no barrier: the 2 MB x NB loads / stores
  // of a lane are independent, so their latency overlaps (the staged, row-at-a-time version spent
  // 42 % of the warp samples waiting on its dependent load -> add -> shuffle chains,
  // profiles/r1_bench_top_kernels_ncu.json of that revision).
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
      if (mrow) {
        if (mvec2) {  // J even, tables 16-byte aligned: one 16-byte load per table and column pair
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
    if (EX > 0) {  // remainder columns: sum the four kk shares, lane c = e finishes column JB + e
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
        if (mrow) {
          const double d = o - __ldg(mrow + jx);
          lsum = fma((wrw ? __ldg(wrw + jx) : 1.0) * d, d, lsum);
        }
      }
    }
    if (a.rowloss) {  // fixed order: the lane's columns, then lanes c = 0..3 -> deterministic
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
      if (c == 0 && valid) a.rowloss[(size_t)row * a.nblk + blk] = lsum;
    }
  }
}

// loss[p] = sum_{m,blk} rowloss[(p*M+m)*nblk + blk] / (M*J)   (fixed order -> deterministic)
static __global__ void cos_loss_finalize2_kernel(const double* __restrict__ rowloss, int P, int M, int J,
                                          int nblk, double* __restrict__ loss) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double acc = 0.0;
  const double* r = rowloss + (size_t)p * M * nblk;
  for (int i = 0; i < M * nblk; ++i) acc += r[i];
  loss[p] = acc / ((double)M * (double)J);
}

}  // namespace fop
