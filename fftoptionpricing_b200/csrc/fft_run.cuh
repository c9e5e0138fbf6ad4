// Host driver of the batched FFT engine (fft_engine.cuh): run_fft<SIGN>(handle, arena, N, batch,
// generator, epilogue) launches the single-CTA kernel (N <= 8192) or the four-step pair.  Shared by
// capi_fft.cu (fop_fft, Carr-Madan, integrateFFT, large-grid FSTS) and capi_samples.cu (the
// caller-sampled entry points).
#pragma once
#include <algorithm>

#include "fft_engine.cuh"
#include "handle.cuh"

namespace fop {
int get_twiddles_shared(fop_handle_t h, int N, const double2** out);  // capi_fft.cu
}

namespace {
using namespace fop;

inline int ilog2(int N) {
  int n = 0;
  while ((1 << n) < N) ++n;
  return n;
}
inline int get_twiddles(fop_handle_t h, int N, const double2** out) { return fop::get_twiddles_shared(h, N, out); }

// ---- generic load / store functors (fop_fft)
struct LoadGen {
  static constexpr bool has_cf = false;
  static size_t aux_bytes(int) { return 0; }
  const cd* data;
  int N;
  __device__ void prepare(void*, int, int) const {}
  __device__ cd operator()(const void*, int, int p, int j) const { return data[(size_t)p * N + j]; }
};
struct StoreEpi {
  cd* data;
  int N;
  __device__ void operator()(int p, int k, cd v) const { data[(size_t)p * N + k] = v; }
};
// runs `batch` transforms of length N through the engine
template <int SIGN, typename Gen, typename Epi>
int run_fft(fop_handle_t h, Arena& ar, int N, int batch, Gen gen, Epi epi) {
  const int n = ilog2(N);
  if (n <= 13) {
    const double2* tw = nullptr;
    FOP_TRY(get_twiddles(h, N, &tw));
    const size_t smem = fft_single_smem(n) + Gen::aux_bytes(fft_single_tpb(n));
    auto kern = fft_single_kernel<SIGN, Gen, Epi>;
    FOP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tpb = fft_single_tpb(n);
    const int groups = (batch + tpb - 1) / tpb;
    const int grid = std::min(groups, h->sm_count * 8);
    {
      ProfScope ps(h, Gen::has_cf ? "carr_madan_single_kernel" : "fft_single_kernel");
      kern<<<grid, fft_single_threads(n), smem, h->stream>>>(n, batch, tw, gen, epi);
    }
    FOP_LAUNCH_CHECK(h);
    return FOP_OK;
  }
  const FourStep f = four_step_plan(n);
  const int N1 = 1 << f.n1, N2 = 1 << f.n2;
  const double2 *tw1 = nullptr, *tw2 = nullptr, *twN = nullptr;
  FOP_TRY(get_twiddles(h, N1, &tw1));
  FOP_TRY(get_twiddles(h, N2, &tw2));
  FOP_TRY(get_twiddles(h, N, &twN));
  // scratch for up to `pb` transforms in flight (<= 512 MiB)
  const int pb = (int)std::max<long long>(1, std::min<long long>(batch, (512ll << 20) / ((long long)N * 16)));
  cd* scratch = ar.alloc<cd>((size_t)pb * N);
  if (!scratch) return fail(h, FOP_OOM, "arena exhausted for four-step scratch");
  auto ka = fft_cols_kernel<SIGN, Gen>;
  auto kb = fft_rows_kernel<SIGN, Epi>;
  const size_t sa = four_step_smem(f.n1, f.ta) + Gen::aux_bytes(1), sb = four_step_smem(f.n2, f.tb);
  FOP_CUDA(h, cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa));
  FOP_CUDA(h, cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
  for (int p0 = 0; p0 < batch; p0 += pb) {
    const int pc = std::min(pb, batch - p0);
    {
      ProfScope ps(h, Gen::has_cf ? "carr_madan_cols_kernel" : "fft_cols_kernel");
      ka<<<dim3(N2 / f.ta, pc), f.ta * (N1 / 16), sa, h->stream>>>(f, p0, tw1, twN, gen, scratch);
    }
    FOP_LAUNCH_CHECK(h);
    {
      ProfScope ps(h, Gen::has_cf ? "carr_madan_rows_kernel" : "fft_rows_kernel");
      kb<<<dim3(N1 / f.tb, pc), f.tb * (N2 / 16), sb, h->stream>>>(f, p0, tw2, scratch, epi);
    }
    FOP_LAUNCH_CHECK(h);
  }
  return FOP_OK;
}
inline size_t fft_scratch_bytes(int N, int batch) {
  if (ilog2(N) <= 13) return 0;
  const long long pb = std::max<long long>(1, std::min<long long>(batch, (512ll << 20) / ((long long)N * 16)));
  return Arena::pad((size_t)pb * N * 16);
}


}  // namespace
