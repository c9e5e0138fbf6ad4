// Handle, error plumbing and buffer staging shared by every C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fftop_b200.h"

struct fop_handle_s {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t smem_per_sm = 0;  // shared memory per SM
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  int last_cos_contraction = 0;  // fop_last_cos_contraction
  // bump arena for per-call device temporaries (grown on demand, reused across calls)
  unsigned char* arena = nullptr;
  size_t arena_cap = 0, arena_used = 0;
  // cached twiddle tables: tw[N] holds e^{-2 pi i k/N}, k < N  (see fft_smem.cuh)
  struct Twiddle { int N; double2* d; };
  std::vector<Twiddle> twiddles;
  // COS strike-grid cache: u_k, V_k, strikes, maturities and the cos/sin basis depend only on
  // (K_desc, S0, T, numU), which a calibration loop keeps fixed while theta changes
  struct CosCache {
    std::vector<double> K, T;
    double S0 = 0;
    double dt = 0;          // > 0: maturities on the grid T_m = n_m dt (n_m stored after T)
    int numU = 0, Jp = 0, JB = 0, KK = 0;
    double* dev = nullptr;  // [uk | vk | x | K | T | basis | vk8 | basis digit planes]
    size_t bytes = 0;
    // fixed-point tables of the int8 contraction (cos_gemm_i8.cuh), built on demand: basis digit planes
    // of price / delta / gamma with their scales amax_g, the coefficient kernel's scale tables
    bool q8_valid[3] = {false, false, false};
    double q8_amax[3] = {0, 0, 0};
    bool q8_vp_valid = false, q8_vt_valid = false;
    double q8_theta_bound = 0;   // 4 + |r| the theta scale table was built for
    std::vector<double> uk_host, vk_host;
  } cos_cache;
  // optional NCCL communicator (capi_comm.cu; opaque here, NCCL is bound at run time)
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 0;
  // side stream of fop_gather_losses_async (the gather overlaps the next step's kernels)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ready = nullptr, comm_done = nullptr;
  bool comm_pending = false;
  // optional per-kernel device timing (fop_profile_*): accumulated by kernel name
  bool profiling = false;
  struct ProfEntry { std::string name; double ms = 0; long long launches = 0; };
  std::vector<ProfEntry> prof;
  struct ProfPending { int entry; cudaEvent_t e0, e1; };
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> event_pool;
};

namespace fop {

inline int fail(fop_handle_t h, int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return status;
}

#define FOP_CUDA(h, call)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fop::fail(h, e__ == cudaErrorMemoryAllocation ? FOP_OOM : FOP_CUDA_ERROR,     \
                       "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,  \
                       __LINE__);                                                          \
  } while (0)

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Per-call arena.  begin() resets the bump pointer; reserve() must be called before any alloc()
// of the call with the total (so the arena never moves while kernels are queued on it).
struct Arena {
  fop_handle_t h;
  explicit Arena(fop_handle_t hh) : h(hh) { h->arena_used = 0; }
  static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
  int reserve(size_t total) {
    total += 4096;
    if (total <= h->arena_cap) return FOP_OK;
    FOP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr;
    h->arena_cap = 0;
    FOP_CUDA(h, cudaMalloc(&h->arena, total));
    h->arena_cap = total;
    return FOP_OK;
  }
  template <typename T>
  T* alloc(size_t count) {
    const size_t b = pad(count * sizeof(T));
    if (h->arena_used + b > h->arena_cap) return nullptr;
    T* p = reinterpret_cast<T*>(h->arena + h->arena_used);
    h->arena_used += b;
    return p;
  }
};

// returns a device pointer for a read-only input (copies host data on the handle's stream)
template <typename T>
inline int stage_in(fop_handle_t h, Arena& ar, const T* src, size_t count, const T** dev) {
  if (is_device_ptr(src)) {
    *dev = src;
    return FOP_OK;
  }
  T* d = ar.alloc<T>(count);
  if (!d) return fail(h, FOP_OOM, "arena exhausted staging %zu bytes", count * sizeof(T));
  FOP_CUDA(h, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  *dev = d;
  return FOP_OK;
}
inline size_t stage_bytes(const void* src, size_t bytes) {
  return is_device_ptr(src) ? 0 : Arena::pad(bytes);
}

#define FOP_TRY(expr)              \
  do {                             \
    int st__ = (expr);             \
    if (st__ != FOP_OK) return st__; \
  } while (0)

// Per-kernel timing scope: { ProfScope ps(h, "kernel_name"); kernel<<<...>>>(...); }
struct ProfScope {
  fop_handle_t h;
  int idx = -1;
  ProfScope(fop_handle_t hh, const char* name) : h(hh) {
    if (!h->profiling) return;
    int e = -1;
    for (size_t i = 0; i < h->prof.size(); ++i)
      if (h->prof[i].name == name) e = (int)i;
    if (e < 0) {
      h->prof.push_back({name, 0.0, 0});
      e = (int)h->prof.size() - 1;
    }
    auto get = [&]() {
      cudaEvent_t ev;
      if (!h->event_pool.empty()) {
        ev = h->event_pool.back();
        h->event_pool.pop_back();
      } else {
        cudaEventCreate(&ev);
      }
      return ev;
    };
    fop_handle_s::ProfPending p{e, get(), get()};
    cudaEventRecord(p.e0, h->stream);
    h->prof_pending.push_back(p);
    idx = (int)h->prof_pending.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof_pending[idx].e1, h->stream);
  }
};

#define FOP_LAUNCH_CHECK(h)                                                               \
  do {                                                                                    \
    (h)->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return fop::fail(h, FOP_CUDA_ERROR, "kernel launch failed: %s (%s:%d)",             \
                       cudaGetErrorString(e__), __FILE__, __LINE__);                      \
  } while (0)

}  // namespace fop
