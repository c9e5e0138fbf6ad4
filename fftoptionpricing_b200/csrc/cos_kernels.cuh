// Fang-Oosterlee COS pricer for sm_100a: characteristic function, COS coefficients and the
// strikes x u contraction fused in one kernel (replaces OptionPricing.h:325-346 +
// fangoost::computeExpectationVectorLevy).
//
//   val[row][j] = sum_k  Re(D_k) V_k cos(u_k (x_j - a))  -  Im(D_k) V_k sin(u_k (x_j - a))
//               = sum_kk C[row][kk] * Bas[kk][j],      kk = 0 .. 2 numU - 1
//   row = (parameter set p, maturity m),  D_k = enhCF(phi_{p,m}(i u_k), i u_k) cp  (k = 0 halved)
//
// The basis Bas[2 numU][J] depends only on the strike list, so it is shared by every row: the
// work is a tall-skinny real fp64 GEMM whose left operand is *generated on chip* and never
// touches HBM.
//
// One CTA owns a tile of R rows:
//   stage 1  all threads evaluate the CF once per (set, u) pair (cf_prepare) and once more per
//            maturity (cf_log_at), and write C[kk][row] into shared memory (full K resident);
//   stage 2  for each block of JB strikes the 8 "k-split" warps each hold the whole R x JB
//            accumulator tile in registers (TR x TJ per thread) and sweep a disjoint eighth of
//            kk; the basis streams L2 -> smem through a cp.async double buffer; partial tiles
//            are tree-reduced through smem and the epilogue (put/call parity, Greek scaling,
//            optional weighted squared error against market quotes) stores coalesced.
// fp64 DFMA is the bound (64 lanes/clk/SM), so the design goal is simply: keep the fp64 pipe
// issuing -- thread tiles are sized so smem wavefronts stay well under DFMA issue cycles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cf_models.cuh"

namespace fop {

struct CosArgs {
  int base, tc, greek, kind;
  int P, M, J, K;   // K = numU
  int Jp, nblk;     // padded strike count Jp = nblk * JB
  int np;
  double r, S0;
  const double* params;   // [P][np]
  const double* T;        // [M]
  const double* uk;       // [K]   u_k = du * k
  const double* vk;       // [K]   V_k * cp * (k == 0 ? 1/2 : 1)
  const double* basis;    // [2K][Jp]
  const double* strikes;  // [J]
  double* out;            // [P*M][J] or nullptr
  const double* mkt;      // [M][J] or nullptr (loss mode)
  const double* w;        // [M][J] or nullptr
  double* rowloss;        // [P*M]  or nullptr
};

constexpr int COS_KC = 64;  // basis rows per smem stage

__host__ __device__ constexpr int cos_rows(int TR, int WR) { return 8 * TR * WR; }

// shared-memory carve-up (bytes), identical on host and device
struct CosSmem {
  size_t c_off, b_off, red_off, consts_off, disc_off, rowacc_off, total;
};
__host__ __device__ inline CosSmem cos_smem_layout(int K, int TR, int TJ, int WR) {
  const int R = cos_rows(TR, WR), RS = R + 2, JB = 4 * TJ, JBP = JB + 2;
  CosSmem s;
  size_t o = 0;
  s.c_off = o;       o += (size_t)2 * K * RS * 8;
  // the reduction tiles alias the basis double buffer (the buffer is idle once a strike block's
  // last chunk has been consumed)
  const size_t bbytes = (size_t)2 * COS_KC * JB * 8, rbytes = (size_t)4 * R * JBP * 8;
  s.b_off = o;
  s.red_off = o;     o += bbytes > rbytes ? bbytes : rbytes;
  s.consts_off = o;  o += (size_t)(R + 1) * sizeof(ModelConsts);
  s.disc_off = o;    o += (size_t)R * 8;
  s.rowacc_off = o;  o += (size_t)R * 8;
  s.total = o;
  return s;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// output map of the ten reference entry points (OptionPricing.h:378,414,439,464,490,515,541)
__device__ __forceinline__ double cos_epilogue(int kind, double val, double disc, double K,
                                               double S0, double r) {
  switch (kind) {
    case 0: return (val - 1.0) * disc * K + S0;
    case 1: return val * disc * K;
    case 2: return val * disc * K / S0 + 1.0;
    case 3: return val * disc * K / S0;
    case 4:
    case 5: return val * disc * K / (S0 * S0);
    case 6: return (val - r) * disc * K;
    default: return val * disc * K;
  }
}

// basis[2k][j] = cos(u_k (x_j - xMin)), basis[2k+1][j] = -sin(u_k (x_j - xMin)); 0 for j >= J
__global__ void cos_basis_kernel(const double* __restrict__ uk, const double* __restrict__ x,
                                 int K, int J, int Jp, double* __restrict__ basis) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * Jp) return;
  const int k = idx / Jp, j = idx - k * Jp;
  double c = 0.0, s = 0.0;
  if (j < J) sincos(uk[k] * (x[j] - x[0]), &s, &c);
  basis[(size_t)(2 * k) * Jp + j] = c;
  basis[(size_t)(2 * k + 1) * Jp + j] = -s;
}

// loss[p] = sum_m rowloss[p*M+m] / (M*J)   (fixed order -> deterministic)
__global__ void cos_loss_finalize_kernel(const double* __restrict__ rowloss, int P, int M, int J,
                                         double* __restrict__ loss) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double acc = 0.0;
  for (int m = 0; m < M; ++m) acc += rowloss[(size_t)p * M + m];
  loss[p] = acc / ((double)M * (double)J);
}

template <int TR, int TJ, int WR>
__global__ void __launch_bounds__(256 * WR, 1) cos_kernel(const CosArgs a) {
  constexpr int R = cos_rows(TR, WR);
  constexpr int RS = R + 2;          // row stride of C in doubles (even: 16 B aligned pairs)
  constexpr int JB = 4 * TJ;         // strikes per block
  constexpr int JBP = JB + 2;        // padded stride of the reduction tiles
  constexpr int NT = 256 * WR;
  constexpr int TJH = TJ / 2;
  static_assert(TJ % 2 == 0, "TJ must be even");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const CosSmem L = cos_smem_layout(a.K, TR, TJ, WR);
  double* Csm = reinterpret_cast<double*>(smem_raw + L.c_off);
  double* Bsm = reinterpret_cast<double*>(smem_raw + L.b_off);
  double* Red = reinterpret_cast<double*>(smem_raw + L.red_off);
  ModelConsts* consts = reinterpret_cast<ModelConsts*>(smem_raw + L.consts_off);
  double* rowdisc = reinterpret_cast<double*>(smem_raw + L.disc_off);
  double* rowacc = reinterpret_cast<double*>(smem_raw + L.rowacc_off);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wk = warp & 7;   // k-split index
  const int wr = warp >> 3;  // row-split index
  const int g = lane >> 2, cg = lane & 3;
  const int K = a.K, M = a.M, J = a.J, KK = 2 * K;
  const long long rows = (long long)a.P * M;
  const long long ntiles = (rows + R - 1) / R;
  const int greek = a.greek;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * R;
    const int nrows = (int)min((long long)R, rows - row0);
    const int p_first = (int)(row0 / M);
    const int p_last = (int)((row0 + nrows - 1) / M);
    const int nsets = p_last - p_first + 1;

    // ---- stage 0: per-set constants, per-row discount
    if (tid < nsets) make_consts(a.base, a.tc, a.params + (size_t)(p_first + tid) * a.np, consts[tid]);
    if (tid < R) {
      rowacc[tid] = 0.0;
      if (tid < nrows) rowdisc[tid] = exp(-a.r * __ldg(a.T + (int)((row0 + tid) % M)));
    }
    __syncthreads();

    // ---- stage 1: characteristic function -> COS coefficients C[kk][row]
    for (int idx = tid; idx < nsets * K; idx += NT) {
      const int s = idx / K, k = idx - s * K;
      const ModelConsts& mc = consts[s];
      const cd u = mk(0.0, __ldg(a.uk + k));
      const double v = __ldg(a.vk + k);
      CfPre pre;
      cf_prepare(a.base, a.tc, mc, u, a.r, pre);
      const long long rbase = (long long)(p_first + s) * M - row0;  // tile row of maturity 0
      const int m_lo = rbase < 0 ? (int)(-rbase) : 0;
      const int m_hi = (int)min((long long)M, (long long)nrows - rbase);
      double* c0 = Csm + (size_t)(2 * k) * RS + rbase;
      double* c1 = c0 + RS;
      int m = m_lo;
      for (; m + 1 < m_hi; m += 2) {  // two maturities per trip: independent chains for ILP
        const double t0 = __ldg(a.T + m), t1 = __ldg(a.T + m + 1);
        const cd l0 = cf_log_at(a.tc, mc, pre, t0);
        const cd l1 = cf_log_at(a.tc, mc, pre, t1);
        const cd d0 = option_transform(greek, cexp(l0), u, a.r);
        const cd d1 = option_transform(greek, cexp(l1), u, a.r);
        c0[m] = d0.x * v;  c1[m] = d0.y * v;
        c0[m + 1] = d1.x * v;  c1[m + 1] = d1.y * v;
      }
      if (m < m_hi) {
        const cd l0 = cf_log_at(a.tc, mc, pre, __ldg(a.T + m));
        const cd d0 = option_transform(greek, cexp(l0), u, a.r);
        c0[m] = d0.x * v;  c1[m] = d0.y * v;
      }
    }
    __syncthreads();

    // ---- stage 2: contraction against the shared basis, one block of JB strikes at a time
    const int nchunks = (KK + COS_KC - 1) / COS_KC;
    const int arow = wr * 8 * TR;  // first tile row of this warp
    for (int blk = 0; blk < a.nblk; ++blk) {
      double acc[TR][TJ];
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0.0;

      const double* bsrc = a.basis + (size_t)blk * JB;
      auto issue_chunk = [&](int c, int buf) {
        const int kk0 = c * COS_KC;
        const int nk = min(COS_KC, KK - kk0);
        constexpr int CPR = JB / 2;  // 16-byte pieces per basis row
        for (int q = tid; q < nk * CPR; q += NT) {
          const int rr = q / CPR, cc = q - rr * CPR;
          cp_async16(Bsm + (size_t)buf * COS_KC * JB + rr * JB + 2 * cc,
                     bsrc + (size_t)(kk0 + rr) * a.Jp + 2 * cc);
        }
        cp_async_commit();
      };
      issue_chunk(0, 0);
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) {
          issue_chunk(c + 1, buf ^ 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
        const int kk0 = c * COS_KC;
        const double* bbuf = Bsm + (size_t)buf * COS_KC * JB;
#pragma unroll 2
        for (int i = 0; i < COS_KC / 8; ++i) {
          const int rr = wk + 8 * i;
          const int kk = kk0 + rr;
          if (kk < KK) {
            double av[TR];
            const double* ap = Csm + (size_t)kk * RS + arow;
            if (TR >= 2) {
#pragma unroll
              for (int t = 0; t < TR / 2; ++t) {
                const double2 v2 = *reinterpret_cast<const double2*>(ap + 16 * t + 2 * g);
                av[2 * t] = v2.x;
                av[2 * t + 1] = v2.y;
              }
            } else {
              av[0] = ap[g];
            }
            const double* bp = bbuf + rr * JB + 2 * cg;
#pragma unroll
            for (int h = 0; h < TJH; ++h) {
              const double2 b2 = *reinterpret_cast<const double2*>(bp + 8 * h);
#pragma unroll
              for (int t = 0; t < TR; ++t) {
                acc[t][2 * h] = fma(av[t], b2.x, acc[t][2 * h]);
                acc[t][2 * h + 1] = fma(av[t], b2.y, acc[t][2 * h + 1]);
              }
            }
          }
        }
        __syncthreads();
      }

      // ---- tree reduction of the 8 k-split partial tiles (per row-split group) through smem.
      // Tile element (row, col): row = arow + rowof(t), col = 2 cg + 8 h + {0,1}.
      auto rowof = [&](int t) { return TR >= 2 ? 16 * (t >> 1) + 2 * g + (t & 1) : g; };
#pragma unroll
      for (int half = 4; half >= 1; half >>= 1) {
        if (wk >= half && wk < 2 * half) {
          double* slot = Red + (size_t)(wk - half) * R * JBP;
#pragma unroll
          for (int t = 0; t < TR; ++t)
#pragma unroll
            for (int h = 0; h < TJH; ++h)
              *reinterpret_cast<double2*>(slot + (size_t)(arow + rowof(t)) * JBP + 2 * cg + 8 * h) =
                  make_double2(acc[t][2 * h], acc[t][2 * h + 1]);
        }
        __syncthreads();
        if (wk < half) {
          const double* slot = Red + (size_t)wk * R * JBP;
#pragma unroll
          for (int t = 0; t < TR; ++t)
#pragma unroll
            for (int h = 0; h < TJH; ++h) {
              const double2 v2 = *reinterpret_cast<const double2*>(
                  slot + (size_t)(arow + rowof(t)) * JBP + 2 * cg + 8 * h);
              acc[t][2 * h] += v2.x;
              acc[t][2 * h + 1] += v2.y;
            }
        }
        __syncthreads();
      }
      if (wk == 0) {
        double* slot = Red;  // final tile
#pragma unroll
        for (int t = 0; t < TR; ++t)
#pragma unroll
          for (int h = 0; h < TJH; ++h)
            *reinterpret_cast<double2*>(slot + (size_t)(arow + rowof(t)) * JBP + 2 * cg + 8 * h) =
                make_double2(acc[t][2 * h], acc[t][2 * h + 1]);
      }
      __syncthreads();

      // ---- epilogue: one warp per row, lanes along strikes (coalesced stores)
      for (int rr = warp; rr < nrows; rr += NT / 32) {
        const long long row = row0 + rr;
        const int m = (int)(row % M);
        const double disc = rowdisc[rr];
        double lsum = 0.0;
        for (int jj = lane; jj < JB; jj += 32) {
          const int j = blk * JB + jj;
          if (j < J) {
            const double val = Red[(size_t)rr * JBP + jj];
            const double o = cos_epilogue(a.kind, val, disc, __ldg(a.strikes + j), a.S0, a.r);
            if (a.out) a.out[(size_t)row * J + j] = o;
            if (a.mkt) {
              const double d = o - __ldg(a.mkt + (size_t)m * J + j);
              lsum += (a.w ? __ldg(a.w + (size_t)m * J + j) : 1.0) * d * d;
            }
          }
        }
        if (a.mkt) {
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
          if (lane == 0) rowacc[rr] += lsum;
        }
      }
      __syncthreads();
    }  // strike blocks

    if (a.rowloss && tid < nrows) a.rowloss[row0 + tid] = rowacc[tid];
    __syncthreads();
  }
}

}  // namespace fop
