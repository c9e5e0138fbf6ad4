// COS contraction on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// tcgen05 has no f64 kind, and the fp64 contraction (cos_gemm_tma.cuh, DMMA) is bound by the fp64
// pipe at 31 of 37 TFLOP/s.  The contraction is nevertheless exact-integer friendly:
//
//   val[row][j] = sum_kk C[row][kk] Bas[kk][j],   C[row][kk] = V_k cp * (Re | Im) phi,  |phi| <= 1,  |Bas| <= 1
//
// so both operands are FIXED-POINT numbers with a scale known before the launch.  The magnitude of
// V_k cp (it decays like 1/k) is moved to the basis so that every coefficient uses its full range:
//   C'[row][kk] = (Re | Im) phi                 = qa 2^-45,   qa integer, |qa| <= 2^45   (coefficient kernel)
//   Bas'[kk][j] = Bas[kk][j] V_k cp / amax      = qb 2^-46,   qb integer, |qb| <= 2^46   (once per strike grid)
//   val = amax sum_kk C' Bas',   amax = max_k |V_k cp|
// Each integer is split into six signed base-256 digits (int8): q = sum_s d_s 256^s.  The product of
// two digit planes is an EXACT int8 x int8 -> int32 tensor-core GEMM (products <= 2^14, at most six
// plane pairs per accumulator: < 2^26 for the usual 512 terms, and below 2^31 for any digits up to
// 16 384 terms = numU 8 192, the limit the host enforces), and
//   qa qb = sum_{s,t} a_s b_t 256^{s+t}.
// The 26 plane pairs with s + t >= 4 are kept, accumulated per level L = s + t - 4 = 0..6 in its own
// TMEM columns; the 10 pairs below contribute < 2^-53 of full scale per term.  What remains is the
// rounding of the operands themselves (2^-46 relative to full scale, weighted by |V_k| resp. |phi|):
// ~2e-14 amax K_j on a price, against ~1e-15 for the fp64 kernel and the parity bar of
// 1e-8 + 1e-10 |ref| -- at the int8 tensor rate instead of the fp64 rate.  (This is the
// error-free-transformation idea of Ozaki et al. specialised to operands with a known bound, which
// removes the per-row exponent search.)  A coefficient that is NaN or out of range (|phi| > 2: not a
// characteristic function) flags its row and the row is returned as NaN.
//
// Kernel: persistent CTAs (one per SM, 320 threads), warp-specialised:
//   warp 0    TMA producer: per 128-kk chunk six boxes of A (128 kk x 128 rows, one per digit plane,
//             each into its own ring slot with its own full / empty mbarrier) and one box of B (128 kk x
//             72 strikes x 6 planes, double buffered), 128-byte swizzle.  (64-byte rows were measured
//             first: the copy engine then needs ~3 000 cycles per 77 KiB stage -- twice as many row
//             requests -- and the tensor pipe sat at 41 %, profiles/r2_cos_gemm_i8_ncu.json.)
//   warp 1    one lane issues tcgen05.mma (M 128, K 32).  For coefficient plane s the basis planes
//             max(0, 4 - s) .. 5 lie contiguously in shared memory (72 rows each) and their levels
//             contiguously in TMEM (72 columns each), so two planes are multiplied by ONE instruction
//             of N = 144; an odd last plane (always plane 5) takes N = 80, reading eight zero rows
//             behind plane 5 and adding zeros to the first eight columns of the next level.  Fourteen
//             instructions per 32 kk instead of 26.  tcgen05.commit releases the stage / publishes
//             the accumulators.
//   (more than 72 strikes: blockIdx.y walks blocks of 72 columns, each a pass over the coefficient planes)
//   warps 2-9 epilogue (a warp reads TMEM lanes 32 (warp % 4) .. +31; two warps share a lane quarter
//             and take 36 columns each): tcgen05.ld the seven level accumulators, write zeros back
//             (tcgen05.st: every MMA accumulates, no first-instruction special case), recombine in
//             integer (three groups below 2^48, exact for any int32 accumulators), convert -- then
//             hand TMEM back to the tensor
//             core and only afterwards run the price map + squared-error loss of the fp64 kernel from
//             the 36 doubles held in registers (measured: the tile period fell from 29 000 to
//             ~25 000 cycles against 21 000 with no epilogue at all).
// TMEM: 7 levels x 72 columns = 504 (+8 overlap) of the 512 columns.
#pragma once
#include <cuda.h>

#include "cos_coeff.cuh"
#include "cos_gemm_tma.cuh"

namespace fop {

constexpr int Q8_NQ = 72;                         // strike columns of a tile = TMEM columns per level
constexpr int Q8_LEVELS = 7;                      // kept levels s + t - 4
constexpr int Q8_KC = 128;                        // kk (bytes) per chunk = one 128-byte swizzle span
constexpr int Q8_STB = 2;                         // stages of the basis ring (L2 resident)
constexpr int Q8_ROWS = 128;                      // rows per tile (UMMA M)
constexpr int Q8_A_PLANE = Q8_ROWS * Q8_KC;       // 16 KiB: one ring slot per coefficient plane
constexpr int Q8_B_PLANE = Q8_NQ * Q8_KC;         // 9 KiB (a multiple of the 1024-byte swizzle pattern)
constexpr int Q8_A_BYTES = Q8_SLICES * Q8_A_PLANE;
constexpr int Q8_B_BYTES = Q8_SLICES * Q8_B_PLANE;
constexpr int Q8_B_PAD = 8 * Q8_KC;               // eight zero rows behind basis plane 5
constexpr int Q8_B_STAGE = Q8_B_BYTES + Q8_B_PAD; // 55 KiB
constexpr int Q8_RING = Q8_A_BYTES + Q8_STB * Q8_B_STAGE;  // 206 KiB
constexpr int Q8_EPI_WARPS = 8;
constexpr int Q8_THREADS = 64 + 32 * Q8_EPI_WARPS;
constexpr int Q8_TMEM_COLS = 512;
constexpr double Q8_A_SCALE = 35184372088832.0;    // 2^45
constexpr double Q8_B_SCALE = 70368744177664.0;    // 2^46

struct CosGemmI8Args {
  int kind;
  int M, J, KK;
  int nblk;                     // strike blocks of 72 columns (blockIdx.y)
  long long rows;
  double r, S0;
  double amax;                  // val = amax 2^-91 sum qa qb
  const double* T;              // [M]
  const double* strikes;        // [J]
  double* out;                  // [rows][J] or nullptr
  int vec2;                     // prices may be stored as 16-byte pairs (J even, out 16-byte aligned)
  const double* mkt;            // [M][J] or nullptr
  const double* w;              // [M][J] or nullptr
  double* rowloss;              // [rows][nblk][2] (one partial sum per block and column half) or nullptr
  const unsigned char* rowflag; // [rows]: row holds a non-representable coefficient -> NaN
};
// ring + alignment slack + market / weight tables of one strike block [2][M][72] + strikes + discounts
__host__ inline size_t cos_i8_smem(int M) {
  return (size_t)Q8_RING + 1024 + ((size_t)2 * M * Q8_NQ + Q8_NQ + M) * 8;
}

// basis digit planes B8[6][J8][KK] (K-major: one row of KK bytes per strike; J8 = strike count padded
// to whole 72-column blocks) of transform g (0 price, 1 delta, 2 gamma; OptionPricing.h:44-73): the
// transform multiplies phi(i u_k) by c_k = 1, i u_k, -(u_k^2 + i u_k), a factor that does not depend on
// the parameter set, so it lives HERE and the coefficient planes are phi itself:
//   Re((a + i b) z) = a Re z - b Im z,  z = c_k e^{i u_k (x_j - x_0)}  ->  rows 2k, 2k + 1 = Re z, -Im z,
// times V_k cp / amax_g (scaling note above).
static __global__ void cos_basis_q8_kernel(const double* __restrict__ uk, const double* __restrict__ vk,
                                           const double* __restrict__ x, int g, double inv_amax, int numU,
                                           int KK, int J, int J8, unsigned char* __restrict__ B8) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= J8 * KK) return;
  const int j = idx / KK, kk = idx - j * KK;
  const int k = kk >> 1;
  double v = 0.0;
  if (j < J && k < numU) {
    const double u = uk[k];
    double sn, cs;
    sincos(u * (x[j] - x[0]), &sn, &cs);
    const double fr = g == 0 ? 1.0 : g == 1 ? 0.0 : -u * u;
    const double fi = g == 0 ? 0.0 : g == 1 ? u : -u;
    const double zr = fr * cs - fi * sn, zi = fr * sn + fi * cs;
    v = ((kk & 1) ? -zi : zr) * (vk[k] * inv_amax);
  }
  unsigned lo, hi;
  q8_digits(v * Q8_B_SCALE, lo, hi);   // |v| <= 1: always representable
  const unsigned long long u8 = lo | ((unsigned long long)hi << 32);
#pragma unroll
  for (int s = 0; s < Q8_SLICES; ++s) B8[((size_t)s * J8 + j) * KK + kk] = (unsigned char)(u8 >> (8 * s));
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          tm_smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tm_smem_u32(bar))
      : "memory");
}
// bounded wait: a protocol error traps instead of hanging the device
__device__ __forceinline__ void q8_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = tm_smem_u32(b);
  for (unsigned spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tm_smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t q8_smem_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3ffff) >> 4);  // start address
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t q8_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(Q8_ROWS >> 4) << 24);
}
__device__ __forceinline__ void q8_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void q8_tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void q8_tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void q8_tmem_ld4(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void q8_tmem_zero16(uint32_t taddr) {
  const int z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(
          taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void q8_tmem_zero8(uint32_t taddr) {
  const int z = 0;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
               : "memory");
}
__device__ __forceinline__ void q8_tmem_zero4(uint32_t taddr) {
  const int z = 0;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
// exact conversion of an integer |x| < 2^51 (one integer add + one DADD instead of I2F.F64.S64)
__device__ __forceinline__ double q8_ll2d(long long x) {
  return __longlong_as_double(x + 0x4338000000000000ll) - 6755399441055744.0;
}

static __global__ void __launch_bounds__(Q8_THREADS, 1)
    cos_gemm_i8_kernel(const CosGemmI8Args a, const __grid_constant__ CUtensorMap tmA,
                       const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_fullA[Q8_SLICES], s_emptyA[Q8_SLICES], s_fullB[Q8_STB], s_emptyB[Q8_STB], s_tfull,
      s_tempty;
  __shared__ uint32_t s_tmem;
  unsigned char* ring = smem_dyn + ((1024u - (tm_smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* ringB = ring + (size_t)Q8_A_BYTES;
  double* s_mkt = reinterpret_cast<double*>(ring + (size_t)Q8_RING);
  double* s_w = s_mkt + (size_t)a.M * Q8_NQ;
  double* s_strike = s_w + (size_t)a.M * Q8_NQ;
  double* s_rs = s_strike + Q8_NQ;
  const int blk = blockIdx.y;
  const int jbase = blk * Q8_NQ;
  const int jvalid = min(Q8_NQ, a.J - jbase);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nch = a.KK / Q8_KC;
  const CosEpi epi = cos_epi_consts(a.kind, a.S0, a.r);
  const long long ntiles = (a.rows + Q8_ROWS - 1) / Q8_ROWS;
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_ch = my_tiles * nch;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < Q8_SLICES; ++s) {
      tm_mbar_init(&s_fullA[s], 1);
      tm_mbar_init(&s_emptyA[s], 1);
    }
#pragma unroll
    for (int s = 0; s < Q8_STB; ++s) {
      tm_mbar_init(&s_fullB[s], 1);
      tm_mbar_init(&s_emptyB[s], 1);
    }
    tm_mbar_init(&s_tfull, 1);
    tm_mbar_init(&s_tempty, 32 * Q8_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp owns the tensor-memory allocation (all 512 columns: one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(&s_tmem)),
                 "r"(Q8_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < a.M * Q8_NQ; i += Q8_THREADS) {
    const int m = i / Q8_NQ, jj = i - m * Q8_NQ;
    const bool in = jj < jvalid;
    s_mkt[i] = (in && a.mkt) ? __ldg(a.mkt + (size_t)m * a.J + jbase + jj) : 0.0;
    s_w[i] = in ? (a.w ? __ldg(a.w + (size_t)m * a.J + jbase + jj) : 1.0) : 0.0;
  }
  if (tid < Q8_NQ) s_strike[tid] = tid < jvalid ? __ldg(a.strikes + jbase + tid) : 0.0;
  if (tid < a.M) s_rs[tid] = exp(-a.r * __ldg(a.T + tid)) * epi.inv;
  // the eight zero rows behind basis plane 5 of every stage (read by the N = 80 instructions)
  for (int i = tid; i < Q8_STB * (Q8_B_PAD / 4); i += Q8_THREADS) {
    const int s = i / (Q8_B_PAD / 4), o = i - s * (Q8_B_PAD / 4);
    reinterpret_cast<uint32_t*>(ringB + (size_t)s * Q8_B_STAGE + Q8_B_BYTES)[o] = 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      // one ring slot per coefficient plane: plane i of chunk g + 1 is fetched as soon as the tensor
      // core has finished plane i of chunk g (the issue order below is plane-major), so ~5/6 of a chunk
      // (80 KiB) is in flight from HBM; the basis chunks (L2) are double buffered
      for (long long gch = 0; gch < total_ch; ++gch) {
        const int sB = (int)(gch % Q8_STB);
        const long long t = gch / nch;
        const int ch = (int)(gch - t * nch);
        const long long r0 = ((long long)blockIdx.x + t * gridDim.x) * Q8_ROWS;
        if (gch >= Q8_STB) q8_wait(&s_emptyB[sB], (uint32_t)(((gch / Q8_STB) - 1) & 1));
        tm_mbar_expect_tx(&s_fullB[sB], Q8_B_BYTES);
        tma_load_3d(ringB + (size_t)sB * Q8_B_STAGE, &tmB, ch * Q8_KC, jbase, 0, &s_fullB[sB]);
#pragma unroll 1
        for (int i = 0; i < Q8_SLICES; ++i) {
          if (gch >= 1) q8_wait(&s_emptyA[i], (uint32_t)((gch - 1) & 1));
          tm_mbar_expect_tx(&s_fullA[i], Q8_A_PLANE);
          tma_load_3d(ring + (size_t)i * Q8_A_PLANE, &tmA, ch * Q8_KC, (int)r0, i, &s_fullA[i]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      long long gch = 0;
      for (long long t = 0; t < my_tiles; ++t) {
        // the accumulators are zero: initially, and again once the epilogue has drained tile t - 1
        q8_wait(&s_tempty, (uint32_t)(t & 1));
        tc_fence_after();
        for (int ch = 0; ch < nch; ++ch, ++gch) {
          const int sB = (int)(gch % Q8_STB);
          q8_wait(&s_fullB[sB], (uint32_t)((gch / Q8_STB) & 1));
          const uint32_t sa = tm_smem_u32(ring);
          const uint32_t sb = tm_smem_u32(ringB + (size_t)sB * Q8_B_STAGE);
#pragma unroll
          for (int i = 0; i < Q8_SLICES; ++i) {
            // coefficient plane i times basis planes t0 .. 5 -> levels i + t0 - 4 .. i + 1
            q8_wait(&s_fullA[i], (uint32_t)(gch & 1));
            tc_fence_after();
            const int t0 = i >= 4 ? 0 : 4 - i;
            const int cnt = Q8_SLICES - t0;
            const int l0 = i + t0 - 4;
#pragma unroll
            for (int ks = 0; ks < Q8_KC / 32; ++ks) {
              const uint64_t ad = q8_smem_desc(sa + i * Q8_A_PLANE + ks * 32);
#pragma unroll
              for (int p = 0; p < cnt; p += 2) {
                const int n = (p + 1 < cnt) ? 2 * Q8_NQ : 80;
                const uint64_t bd = q8_smem_desc(sb + (t0 + p) * Q8_B_PLANE + ks * 32);
                q8_mma(tmem + (l0 + p) * Q8_NQ, ad, bd, q8_idesc(n), 1u);
              }
            }
            tc_commit(&s_emptyA[i]);  // slot free once these MMAs have read it
          }
          tc_commit(&s_emptyB[sB]);
        }
        tc_commit(&s_tfull);       // accumulators complete
      }
    }
  } else {
    // ===== epilogue (warps 2..9) =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // column half: columns 36 half .. 36 half + 35
    const int rr = q * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + half * 36;
    // val = amax 2^-91 sum_L S_L 256^(L + 4) = amax (lo 2^-59 + mid 2^-35 + top 2^-11)
    const double c_lo = a.amax * 0x1p-59;
    const double c_mid = a.amax * 0x1p-35;
    const double c_top = a.amax * 0x1p-11;
    // zero the accumulators once (every MMA accumulates)
#pragma unroll
    for (int L = 0; L < Q8_LEVELS; ++L) {
      q8_tmem_zero16(taddr + L * Q8_NQ);
      q8_tmem_zero16(taddr + L * Q8_NQ + 16);
      q8_tmem_zero4(taddr + L * Q8_NQ + 32);
    }
    // (columns 504 .. 511 only ever receive the zero products of the top level's N = 80 instruction
    // and are never read)
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    tm_mbar_arrive(&s_tempty);
    for (long long t = 0; t < my_tiles; ++t) {
      const long long row0 = ((long long)blockIdx.x + t * gridDim.x) * Q8_ROWS;
      const long long row = row0 + rr;
      const bool valid = row < a.rows;
      const int m = valid ? (int)(row % a.M) : 0;
      const double rs = s_rs[m];
      const bool bad = valid && a.rowflag && a.rowflag[row] != 0;
      double* orow = a.out && valid ? a.out + (size_t)row * a.J + jbase : nullptr;
      const double* mrow = s_mkt + (size_t)m * Q8_NQ;
      const double* wrow = s_w + (size_t)m * Q8_NQ;
      double lsum = 0.0;
      q8_wait(&s_tfull, (uint32_t)(t & 1));
      tc_fence_after();
      // Phase 1 (holds up the tensor core): drain the accumulators -- 8 + 8 + 8 + 8 + 4 columns, seven
      // levels each -- write zeros back, fold the levels into one double per column.
      double vals[36];
#pragma unroll
      for (int cb = 0; cb < 5; ++cb) {
        int S[Q8_LEVELS][8];
        if (cb < 4) {
#pragma unroll
          for (int L = 0; L < Q8_LEVELS; ++L) q8_tmem_ld8(taddr + L * Q8_NQ + cb * 8, S[L]);
        } else {
#pragma unroll
          for (int L = 0; L < Q8_LEVELS; ++L) q8_tmem_ld4(taddr + L * Q8_NQ + 32, S[L]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cb < 4) {
#pragma unroll
          for (int L = 0; L < Q8_LEVELS; ++L) q8_tmem_zero8(taddr + L * Q8_NQ + cb * 8);
        } else {
#pragma unroll
          for (int L = 0; L < Q8_LEVELS; ++L) q8_tmem_zero4(taddr + L * Q8_NQ + 32);
        }
#pragma unroll
        for (int e = 0; e < (cb < 4 ? 8 : 4); ++e) {
          // sum_L S_L 256^L in three groups (levels 0..2, 3..5, 6) that stay below 2^48 for ANY int32
          // accumulators, so the integer -> double conversion by the 1.5 * 2^52 constant is exact; the
          // wide multiply-adds start from that constant: six integer instructions, three exact DADDs
          long long lo = 0x4338000000000000ll, mid = 0x4338000000000000ll, top = 0x4338000000000000ll;
          lo += (long long)S[2][e] * 65536;
          lo += (long long)S[1][e] * 256;
          lo += (long long)S[0][e];
          mid += (long long)S[5][e] * 65536;
          mid += (long long)S[4][e] * 256;
          mid += (long long)S[3][e];
          top += (long long)S[6][e];
          const double dlo = __longlong_as_double(lo) - 6755399441055744.0;
          const double dmid = __longlong_as_double(mid) - 6755399441055744.0;
          const double dtop = __longlong_as_double(top) - 6755399441055744.0;
          vals[cb * 8 + e] = fma(dtop, c_top, fma(dmid, c_mid, dlo * c_lo));
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      tm_mbar_arrive(&s_tempty);  // every epilogue thread has read and re-zeroed its accumulators
      // Phase 2 (overlaps the next tile's MMAs): price map, store, squared-error loss
      const double nanv = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
      for (int c = 0; c < 36; c += 2) {
        const int j = half * 36 + c;
        double o0 = fma(vals[c] - epi.sub, rs * s_strike[j], epi.add);
        double o1 = fma(vals[c + 1] - epi.sub, rs * s_strike[j + 1], epi.add);
        if (bad) o0 = o1 = nanv;
        if (orow) {
          if (a.vec2) {   // (jvalid is even when J is)
            if (j < jvalid) *reinterpret_cast<double2*>(orow + j) = make_double2(o0, o1);
          } else {
            if (j < jvalid) orow[j] = o0;
            if (j + 1 < jvalid) orow[j + 1] = o1;
          }
        }
        if (a.mkt) {   // weights are zero outside the valid columns
          const double d0 = o0 - mrow[j], d1 = o1 - mrow[j + 1];
          lsum = fma(wrow[j] * d0, d0, lsum);
          lsum = fma(wrow[j + 1] * d1, d1, lsum);
        }
      }
      if (a.rowloss && valid) a.rowloss[(row * a.nblk + blk) * 2 + half] = lsum;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Q8_TMEM_COLS) : "memory");
  }
}

}  // namespace fop
