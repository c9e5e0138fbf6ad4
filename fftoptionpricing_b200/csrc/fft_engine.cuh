// Batched forward/inverse FFT engine on top of fft_smem.cuh.
//
//  * N <= 8192: one transform per group of N/16 threads, several transforms per CTA, everything in
//    shared memory (fft_single_kernel).
//  * N  > 8192 (up to 2^20): four-step N = N1 x N2 with one HBM/L2 round trip of the intermediate
//    (fft_cols_kernel: N2 column transforms of length N1 + inter-step twiddle, written transposed;
//    fft_rows_kernel: N1 row transforms of length N2 + epilogue).
//
// Input is produced by a generator functor   gen(aux, p0, p, j) -> cd  (p = transform, j = natural
// index; aux = Gen::aux_bytes of shared memory filled by the block-wide hook gen.prepare(aux, p0,
// count) for transforms p0 .. p0+count-1), output consumed by an epilogue functor epi(p, k, X_k)
// (k = natural frequency index),
// so the same engine serves fop_fft (load / store) and Carr-Madan (characteristic function with
// damping and Simpson weights in, undamped prices out) without the spectrum ever touching HBM
// in natural order.
#pragma once
#include "fft_smem.cuh"

namespace fop {

// tw[k] = e^{-2 pi i k / N}, exact to < 1 ulp (sincospi of an exactly representable argument)
static __global__ void twiddle_table_kernel(double2* tw, int N) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  double s, c;
  sincospi(-2.0 * (double)k / (double)N, &s, &c);
  tw[k] = make_double2(c, s);
}

// Threads per CTA and transforms per CTA for single-CTA sizes
__host__ __device__ inline int fft_single_tpb(int n) {  // transforms per block
  const int T = (1 << n) / 16;
  return T >= 256 ? 1 : 256 / T;
}
__host__ __device__ inline int fft_single_threads(int n) {
  const int T = (1 << n) / 16;
  return T >= 256 ? T : 256;
}
__host__ __device__ inline size_t fft_single_smem(int n) {
  return (size_t)fft_single_tpb(n) * fft_padded_size(1 << n) * sizeof(cd);
}

// LOGN > 0 fixes the transform length at compile time (index algebra of the passes and of the
// digit reversal folds to constants); LOGN = 0 takes it from the argument.
template <int SIGN, typename Gen, typename Epi, int LOGN = 0>
__global__ void __launch_bounds__(512) fft_single_kernel(int n_arg, int batch, const double2* __restrict__ tw, Gen gen,
                                  Epi epi) {
  extern __shared__ __align__(16) unsigned char fft_smem_raw[];
  const int n = LOGN > 0 ? LOGN : n_arg;
  const int N = 1 << n, T = N >> 4;
  const int tpb = fft_single_tpb(n);
  const int psz = fft_padded_size(N);
  const int slot = threadIdx.x / T, b = threadIdx.x - slot * T;
  cd* sm0 = reinterpret_cast<cd*>(fft_smem_raw);
  cd* sm = sm0 + (size_t)slot * psz;
  void* aux = fft_smem_raw + (size_t)tpb * psz * sizeof(cd);
  const int np = fft_num_passes(n);
  for (int p0 = blockIdx.x * tpb; p0 < batch; p0 += gridDim.x * tpb) {
    const int cnt = min(tpb, batch - p0);
    const bool live = slot < cnt;
    __syncthreads();
    gen.prepare(aux, p0, cnt);  // block-wide hook (per-set constants into aux shared memory)
    // cooperative generation, natural order (consecutive threads -> consecutive j); UN independent
    // elements per trip: two straight-line characteristic-function evaluations interleave, a
    // load-only generator gets eight so that its L2 round trips overlap
    constexpr int UN = Gen::has_cf ? 2 : 8;
#pragma unroll 1
    for (int idx = threadIdx.x; idx < cnt * N; idx += UN * blockDim.x) {
      cd v[UN];
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int i = idx + q * blockDim.x;
        const int ii = i < cnt * N ? i : idx;
        v[q] = gen(aux, p0, p0 + (ii >> n), ii & (N - 1));
      }
#pragma unroll
      for (int q = 0; q < UN; ++q) {
        const int i = idx + q * blockDim.x;
        if (i < cnt * N) sm0[(size_t)(i >> n) * psz + fft_pad(i & (N - 1))] = v[q];
      }
    }
    __syncthreads();
    if (LOGN > 0) {
#pragma unroll
      for (int ps = 0; ps < (LOGN + 3) / 4; ++ps) {
        if (live) fft_dif_pass<SIGN>(sm, b, n, ps, tw);
        __syncthreads();
      }
    } else {
      for (int ps = 0; ps < np; ++ps) {
        if (live) fft_dif_pass<SIGN>(sm, b, n, ps, tw);
        __syncthreads();
      }
    }
#pragma unroll 4
    for (int idx = threadIdx.x; idx < cnt * N; idx += blockDim.x) {
      const int sl = idx >> n, k = idx & (N - 1);
      epi(p0 + sl, k, sm0[(size_t)sl * psz + fft_pad(fft_pos(k, n))]);
    }
  }
}

// ---- four-step ---------------------------------------------------------------------------
struct FourStep {
  int n, n1, n2;  // N = 2^n = N1 * N2, N1 = 2^n1 (columns transform length), N2 = 2^n2
  int ta, tb;     // columns per CTA in step A, rows per CTA in step B
};
__host__ __device__ inline FourStep four_step_plan(int n) {
  FourStep f;
  f.n = n;
  f.n1 = (n + 1) / 2;
  f.n2 = n - f.n1;
  f.ta = (1 << f.n1) <= 256 ? 16 : 8;
  f.tb = (1 << f.n2) <= 256 ? 16 : 8;
  return f;
}
__host__ __device__ inline size_t four_step_smem(int nsub, int t) {
  return (size_t)t * fft_padded_size(1 << nsub) * sizeof(cd);
}

// Step A: grid.x = (N2 / ta) column tiles, grid.y = transforms of this launch.
// scratch[p][k1][j2] = W_N^{j2 k1} * sum_{j1} x[j1 N2 + j2] W_N1^{j1 k1}
template <int SIGN, typename Gen>
__global__ void __launch_bounds__(512) fft_cols_kernel(FourStep f, int p_base, const double2* __restrict__ tw1,
                                const double2* __restrict__ twN, Gen gen, cd* __restrict__ scratch) {
  extern __shared__ __align__(16) unsigned char fft_smem_raw[];
  const int N1 = 1 << f.n1, N2 = 1 << f.n2, T = N1 >> 4;
  const int c = threadIdx.x / T, b = threadIdx.x - c * T;
  const int p = p_base + blockIdx.y;
  cd* sm = reinterpret_cast<cd*>(fft_smem_raw) + (size_t)c * fft_padded_size(N1);
  const int np = fft_num_passes(f.n1);
  const int psz = fft_padded_size(N1);
  cd* smA = reinterpret_cast<cd*>(fft_smem_raw);
  void* aux = fft_smem_raw + (size_t)f.ta * psz * sizeof(cd);
  gen.prepare(aux, p, 1);
  // two independent elements per trip (N1 * ta is a multiple of 2 * blockDim)
#pragma unroll 1
  for (int idx = threadIdx.x; idx < N1 * f.ta; idx += 2 * blockDim.x) {
    const int idx1 = idx + blockDim.x;
    const int c0 = idx % f.ta, r0 = idx / f.ta, c1 = idx1 % f.ta, r1 = idx1 / f.ta;
    const cd v0 = gen(aux, p, p, r0 * N2 + blockIdx.x * f.ta + c0);
    const cd v1 = gen(aux, p, p, r1 * N2 + blockIdx.x * f.ta + c1);
    smA[(size_t)c0 * psz + fft_pad(r0)] = v0;
    smA[(size_t)c1 * psz + fft_pad(r1)] = v1;
  }
  __syncthreads();
  for (int ps = 0; ps < np; ++ps) {
    fft_dif_pass<SIGN>(sm, b, f.n1, ps, tw1);
    __syncthreads();
  }
  // transposed, twiddled store: consecutive threads -> consecutive j2 (ta * 16 B segments)
  cd* dst = scratch + (size_t)blockIdx.y * ((size_t)N1 * N2);
  const cd* sm0 = reinterpret_cast<const cd*>(fft_smem_raw);
  for (int idx = threadIdx.x; idx < N1 * f.ta; idx += blockDim.x) {
    const int cc = idx % f.ta, k1 = idx / f.ta;
    const int jj = blockIdx.x * f.ta + cc;
    const cd v = sm0[(size_t)cc * fft_padded_size(N1) + fft_pad(fft_pos(k1, f.n1))];
    const cd w = tw_load<SIGN>(twN, jj * k1);
    dst[(size_t)k1 * N2 + jj] = v * w;
  }
}

// Step B: grid.x = (N1 / tb) row tiles, grid.y = transforms.  X[k1 + N1 k2] = sum_j2 W_N2^{j2 k2} s[k1][j2]
template <int SIGN, typename Epi>
__global__ void __launch_bounds__(512) fft_rows_kernel(FourStep f, int p_base, const double2* __restrict__ tw2,
                                const cd* __restrict__ scratch, Epi epi) {
  extern __shared__ __align__(16) unsigned char fft_smem_raw[];
  const int N1 = 1 << f.n1, N2 = 1 << f.n2, T = N2 >> 4;
  const int c = threadIdx.x / T, b = threadIdx.x - c * T;
  const int p = p_base + blockIdx.y;
  cd* sm0 = reinterpret_cast<cd*>(fft_smem_raw);
  cd* sm = sm0 + (size_t)c * fft_padded_size(N2);
  const cd* src = scratch + (size_t)blockIdx.y * ((size_t)N1 * N2) + (size_t)blockIdx.x * f.tb * N2;
  // coalesced load of tb contiguous rows
  for (int idx = threadIdx.x; idx < f.tb * N2; idx += blockDim.x) {
    const int cc = idx / N2, j2 = idx - cc * N2;
    sm0[(size_t)cc * fft_padded_size(N2) + fft_pad(j2)] = src[idx];
  }
  __syncthreads();
  const int np = fft_num_passes(f.n2);
  for (int ps = 0; ps < np; ++ps) {
    fft_dif_pass<SIGN>(sm, b, f.n2, ps, tw2);
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < f.tb * N2; idx += blockDim.x) {
    const int cc = idx % f.tb, k2 = idx / f.tb;
    const int k1 = blockIdx.x * f.tb + cc;
    epi(p, k1 + N1 * k2, sm0[(size_t)cc * fft_padded_size(N2) + fft_pad(fft_pos(k2, f.n2))]);
  }
}

}  // namespace fop
