// C ABI: handle life cycle, grid helpers, fp64 peak micro-benchmarks (include/fftop_b200.h)
#include <cmath>
#include <new>

#include "cf_models.cuh"
#include "handle.cuh"

namespace {
using namespace fop;

// ---- fp64 throughput probes (roofline denominators; MEASURED_PEAKS.json has no fp64 figure)
// 8 independent DFMA chains per thread, 1024 threads/SM resident.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
         a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;
}

// mode 0: mma.sync m8n8k4 f64 (4 independent accumulator tiles per warp)
// mode 1: mma.sync m16n8k8 f64
// mode 2: even warps DMMA m8n8k4, odd warps DFMA (do the two pipes overlap?)
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed,
                                                        int mode) {
  const int warp = threadIdx.x >> 5;
  double s = 0.0;
  if (mode == 0 || (mode == 2 && (warp & 1) == 0)) {
    double c[4][2];
    for (int t = 0; t < 4; ++t) { c[t][0] = seed + t; c[t][1] = seed - t; }
    const double a = 1.0000001, b = 0.9999999;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        asm volatile(
            "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
            : "+d"(c[t][0]), "+d"(c[t][1])
            : "d"(a), "d"(b));
    }
    for (int t = 0; t < 4; ++t) s += c[t][0] + c[t][1];
  } else if (mode == 1) {
    double c[4][4];
    for (int t = 0; t < 4; ++t)
      for (int q = 0; q < 4; ++q) c[t][q] = seed + t + q;
    const double a = 1.0000001, b = 0.9999999;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        asm volatile(
            "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4}, "
            "{%5,%5}, {%0,%1,%2,%3};\n"
            : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
            : "d"(a), "d"(b));
    }
    for (int t = 0; t < 4; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
  } else {  // mode 2, odd warps: plain DFMA
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
           a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000000001, cc = 1e-9;
    for (int i = 0; i < 4 * iters; ++i) {  // 4x: same FMA count per warp as a DMMA warp
      a0 = fma(a0, m, cc); a1 = fma(a1, m, cc); a2 = fma(a2, m, cc); a3 = fma(a3, m, cc);
      a4 = fma(a4, m, cc); a5 = fma(a5, m, cc); a6 = fma(a6, m, cc); a7 = fma(a7, m, cc);
    }
    s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  }
  if (s == 123.456) out[0] = s;
}

template <typename F>
int time_kernel(fop_handle_t h, F launch, float* ms_best) {
  cudaEvent_t e0, e1;
  FOP_CUDA(h, cudaEventCreate(&e0));
  FOP_CUDA(h, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    FOP_CUDA(h, cudaEventRecord(e0, h->stream));
    launch();
    FOP_CUDA(h, cudaEventRecord(e1, h->stream));
    FOP_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    FOP_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_best = best;
  return FOP_OK;
}

}  // namespace

extern "C" {

int fop_abi_version(void) { return FOP_ABI_VERSION; }

const char* fop_status_string(int status) {
  switch (status) {
    case FOP_OK: return "ok";
    case FOP_INVALID_ARG: return "invalid argument";
    case FOP_CUDA_ERROR: return "CUDA error";
    case FOP_OOM: return "out of device memory";
    case FOP_UNSUPPORTED: return "unsupported configuration";
    case FOP_NCCL_ERROR: return "NCCL error";
    default: return "unknown status";
  }
}

int fop_model_num_params(fop_model_t model) { return fop::num_params(model.base, model.time_change); }

int fop_create(fop_handle_t* handle, int device_ordinal) {
  if (!handle) return FOP_INVALID_ARG;
  *handle = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_ordinal < 0 ||
      device_ordinal >= count) {
    cudaGetLastError();
    return FOP_CUDA_ERROR;  // no CPU fallback by design
  }
  fop_handle_s* h = new (std::nothrow) fop_handle_s();
  if (!h) return FOP_OOM;
  h->device = device_ordinal;
  cudaDeviceProp prop;
  if (cudaSetDevice(device_ordinal) != cudaSuccess ||
      cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete h;
    return FOP_CUDA_ERROR;
  }
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->smem_per_sm = prop.sharedMemPerMultiprocessor;
  *handle = h;
  return FOP_OK;
}

int fop_destroy(fop_handle_t h) {
  if (!h) return FOP_INVALID_ARG;
  fop_comm_destroy(h);
  cudaSetDevice(h->device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  if (h->arena) cudaFree(h->arena);
  if (h->cos_cache.dev) cudaFree(h->cos_cache.dev);
  for (auto& t : h->twiddles) cudaFree(t.d);
  for (auto& p : h->prof_pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
  for (auto& e : h->event_pool) cudaEventDestroy(e);
  delete h;
  return FOP_OK;
}

const char* fop_last_error(fop_handle_t h) { return h ? h->err.c_str() : "null handle"; }
void* fop_get_stream(fop_handle_t h) { return h ? (void*)h->stream : nullptr; }
long long fop_launch_count(fop_handle_t h) { return h ? h->launches : 0; }
int fop_last_cos_contraction(fop_handle_t h) { return h ? h->last_cos_contraction : 0; }

// ---- per-kernel device timing (used by bench.py for the live roofline) ------------------------
int fop_profile_enable(fop_handle_t h, int enable) {
  if (!h) return FOP_INVALID_ARG;
  h->profiling = enable != 0;
  return FOP_OK;
}
// Drains pending events (synchronises the stream), then reports accumulated milliseconds and
// launch counts per kernel name; names are written NUL-separated into names_buf.  Resets.
int fop_profile_read(fop_handle_t h, char* names_buf, int buf_len, double* ms_out,
                     long long* launches_out, int max_entries, int* n_out) {
  if (!h || !n_out) return FOP_INVALID_ARG;
  FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (auto& p : h->prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      h->prof[p.entry].ms += ms;
      h->prof[p.entry].launches += 1;
    }
    h->event_pool.push_back(p.e0);
    h->event_pool.push_back(p.e1);
  }
  h->prof_pending.clear();
  int n = 0, off = 0;
  for (auto& e : h->prof) {
    if (n >= max_entries) break;
    const int len = (int)e.name.size() + 1;
    if (names_buf && off + len <= buf_len) {
      memcpy(names_buf + off, e.name.c_str(), len);
      off += len;
    }
    if (ms_out) ms_out[n] = e.ms;
    if (launches_out) launches_out[n] = e.launches;
    ++n;
  }
  *n_out = n;
  h->prof.clear();
  return FOP_OK;
}

int fop_synchronize(fop_handle_t h) {
  if (!h) return FOP_INVALID_ARG;
  FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  return FOP_OK;
}

// getCarrMadanStrikes, OptionPricing.h:131-138 (b = pi/ada :31-33, lambda = 2b/N :40-42)
int fop_carr_madan_strikes(int N, double ada, double S0, double* K_out) {
  if (N < 1 || !K_out) return FOP_INVALID_ARG;
  const double b = M_PI / ada, lambda = 2.0 * b / N;
  for (int i = 0; i < N; ++i) K_out[i] = S0 * std::exp(-b + lambda * i);
  return FOP_OK;
}
// getFangOostStrike, OptionPricing.h:98-105
int fop_fang_oost_strikes(double xMin, double xMax, double S0, int numX, double* K_out) {
  if (numX < 2 || !K_out) return FOP_INVALID_ARG;
  const double dx = (xMax - xMin) / (numX - 1.0);
  for (int i = 0; i < numX; ++i) K_out[i] = S0 * std::exp(-(xMin + dx * i));
  return FOP_OK;
}
// getStrikeUnderlying, OptionPricing.h:84-90
int fop_strike_underlying(double xMin, double xMax, double K, int numX, double* S_out) {
  if (numX < 2 || !S_out) return FOP_INVALID_ARG;
  const double dx = (xMax - xMin) / (numX - 1.0);
  for (int i = 0; i < numX; ++i) S_out[i] = K * std::exp(xMin + dx * i);
  return FOP_OK;
}

int fop_measure_fp64_peak(fop_handle_t h, double* tflops_out) {
  if (!h || !tflops_out) return FOP_INVALID_ARG;
  FOP_CUDA(h, cudaSetDevice(h->device));
  Arena ar(h);
  FOP_TRY(ar.reserve(256));
  double* d = ar.alloc<double>(1);
  const int iters = 1 << 15, grid = h->sm_count * 4;
  float ms = 0;
  FOP_TRY(time_kernel(h, [&] { dfma_peak_kernel<<<grid, 256, 0, h->stream>>>(d, iters, 1.0); }, &ms));
  h->launches += 5;
  *tflops_out = 2.0 * 8.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
  return FOP_OK;
}

// out[0] DFMA, out[1] DMMA m8n8k4, out[2] DMMA m16n8k8, out[3] mixed (sum of both halves) TFLOP/s
int fop_microbench_fp64(fop_handle_t h, double* out4) {
  if (!h || !out4) return FOP_INVALID_ARG;
  FOP_TRY(fop_measure_fp64_peak(h, &out4[0]));
  Arena ar(h);
  FOP_TRY(ar.reserve(256));
  double* d = ar.alloc<double>(1);
  const int iters = 1 << 13, grid = h->sm_count * 4;
  float ms = 0;
  FOP_TRY(time_kernel(h, [&] { dmma_peak_kernel<<<grid, 256, 0, h->stream>>>(d, iters, 1.0, 0); }, &ms));
  // per warp per iter: 4 mma x (8*8*4) FMA
  out4[1] = 2.0 * 4 * 256.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
  FOP_TRY(time_kernel(h, [&] { dmma_peak_kernel<<<grid, 256, 0, h->stream>>>(d, iters, 1.0, 1); }, &ms));
  out4[2] = 2.0 * 4 * 1024.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;
  FOP_TRY(time_kernel(h, [&] { dmma_peak_kernel<<<grid, 256, 0, h->stream>>>(d, iters, 1.0, 2); }, &ms));
  // 4 warps DMMA + 4 warps DFMA, 1024 FMA per warp per iteration each
  out4[3] = 2.0 * (8 * 1024.0) * iters * grid / (ms * 1e-3) / 1e12;
  h->launches += 15;
  return FOP_OK;
}

}  // extern "C"
