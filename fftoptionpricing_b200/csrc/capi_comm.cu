// C ABI: fop_comm_* / fop_gather_losses -- the one collective of the path: parameter
// sets are sharded over ranks with no exchange during pricing; afterwards the per-set losses are
// all-gathered over NCCL (NVLink / NVSwitch) on the handle's stream, straight from the buffer the
// loss kernel wrote.  One process per GPU; rank 0 creates the 128-byte NCCL id and the host program
// distributes it (file, socket, MPI, torch store ...).  A Python program that already runs
// torch.distributed can use its all_gather instead (bench.py does).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy a host framework has already
// loaded, if any), so the library has no link-time NCCL dependency and single-GPU users never
// touch it.
#include <dlfcn.h>

#include <cstring>

#include "handle.cuh"

namespace {
using namespace fop;

// the subset of the NCCL API used here (ABI stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // 0 = ncclSuccess
enum { ncclFloat64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib || api.ok) return api;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
  return api;
}
int nccl_fail(fop_handle_t h, const char* what, ncclResult_t r) {
  NcclApi& a = nccl();
  return fail(h, FOP_NCCL_ERROR, "%s failed: %s", what,
              a.GetErrorString ? a.GetErrorString(r) : "NCCL error");
}

}  // namespace

extern "C" {

int fop_comm_unique_id(char* id_out128) {
  if (!id_out128) return FOP_INVALID_ARG;
  NcclApi& a = nccl();
  if (!a.ok) return FOP_NCCL_ERROR;
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != 0) return FOP_NCCL_ERROR;
  std::memcpy(id_out128, id.internal, 128);
  return FOP_OK;
}

int fop_comm_init(fop_handle_t h, const char* id128, int rank, int world) {
  if (!h) return FOP_INVALID_ARG;
  if (!id128 || world < 1 || rank < 0 || rank >= world)
    return fail(h, FOP_INVALID_ARG, "fop_comm_init: need id, 0 <= rank < world");
  if (h->comm) return fail(h, FOP_INVALID_ARG, "fop_comm_init: communicator already initialised");
  NcclApi& a = nccl();
  if (!a.ok) return fail(h, FOP_NCCL_ERROR, "libnccl.so.2 could not be loaded");
  FOP_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  const ncclResult_t r = a.CommInitRank(&c, world, id, rank);
  if (r != 0) return nccl_fail(h, "ncclCommInitRank", r);
  h->comm = c;
  h->comm_rank = rank;
  h->comm_world = world;
  return FOP_OK;
}

int fop_gather_losses(fop_handle_t h, const double* loss_local, int count, double* loss_all) {
  if (!h) return FOP_INVALID_ARG;
  if (!h->comm) return fail(h, FOP_INVALID_ARG, "fop_gather_losses: call fop_comm_init first");
  if (!loss_local || !loss_all || count < 0) return fail(h, FOP_INVALID_ARG, "null pointer / negative count");
  if (count == 0) return FOP_OK;
  NcclApi& a = nccl();
  FOP_CUDA(h, cudaSetDevice(h->device));
  const int world = h->comm_world;
  const bool in_host = !is_device_ptr(loss_local), out_host = !is_device_ptr(loss_all);
  Arena ar(h);
  FOP_TRY(ar.reserve((in_host ? Arena::pad((size_t)count * 8) : 0) +
                     (out_host ? Arena::pad((size_t)count * world * 8) : 0)));
  const double* src = loss_local;
  FOP_TRY(stage_in(h, ar, loss_local, (size_t)count, &src));
  double* dst = out_host ? ar.alloc<double>((size_t)count * world) : loss_all;
  const ncclResult_t r = a.AllGather(src, dst, (size_t)count, ncclFloat64,
                                     static_cast<ncclComm_t>(h->comm), h->stream);
  if (r != 0) return nccl_fail(h, "ncclAllGather", r);
  if (out_host)
    FOP_CUDA(h, cudaMemcpyAsync(loss_all, dst, (size_t)count * world * 8, cudaMemcpyDeviceToHost, h->stream));
  // device output: enqueue only (include/fftop_b200.h conventions); host output: return when it is there
  if (out_host) FOP_CUDA(h, cudaStreamSynchronize(h->stream));
  return FOP_OK;
}

int fop_gather_losses_async(fop_handle_t h, const double* loss_local, int count, double* loss_all) {
  if (!h) return FOP_INVALID_ARG;
  if (!h->comm) return fail(h, FOP_INVALID_ARG, "fop_gather_losses_async: call fop_comm_init first");
  if (!loss_local || !loss_all || count < 0) return fail(h, FOP_INVALID_ARG, "null pointer / negative count");
  if (!is_device_ptr(loss_local) || !is_device_ptr(loss_all))
    return fail(h, FOP_INVALID_ARG, "fop_gather_losses_async takes device pointers (host buffers: fop_gather_losses)");
  if (count == 0) return FOP_OK;
  NcclApi& a = nccl();
  FOP_CUDA(h, cudaSetDevice(h->device));
  if (!h->comm_stream) {
    FOP_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    FOP_CUDA(h, cudaEventCreateWithFlags(&h->comm_ready, cudaEventDisableTiming));
    FOP_CUDA(h, cudaEventCreateWithFlags(&h->comm_done, cudaEventDisableTiming));
  }
  // Kernels enqueued after this call may overwrite the buffers of the PREVIOUS async gather (two
  // buffer pairs alternate), so the handle's stream first waits for that one -- it has had a whole
  // step to finish.  The new gather starts when everything enqueued so far on the handle's stream
  // (the loss kernel) is done, and runs beside whatever the caller enqueues next.
  if (h->comm_pending) FOP_CUDA(h, cudaStreamWaitEvent(h->stream, h->comm_done, 0));
  FOP_CUDA(h, cudaEventRecord(h->comm_ready, h->stream));
  FOP_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->comm_ready, 0));
  const ncclResult_t r = a.AllGather(loss_local, loss_all, (size_t)count, ncclFloat64,
                                     static_cast<ncclComm_t>(h->comm), h->comm_stream);
  if (r != 0) return nccl_fail(h, "ncclAllGather", r);
  FOP_CUDA(h, cudaEventRecord(h->comm_done, h->comm_stream));
  h->comm_pending = true;
  return FOP_OK;
}

int fop_comm_wait(fop_handle_t h) {
  if (!h) return FOP_INVALID_ARG;
  if (!h->comm_pending) return FOP_OK;
  FOP_CUDA(h, cudaStreamWaitEvent(h->stream, h->comm_done, 0));
  h->comm_pending = false;
  return FOP_OK;
}

int fop_comm_destroy(fop_handle_t h) {
  if (!h) return FOP_INVALID_ARG;
  if (!h->comm) return FOP_OK;
  NcclApi& a = nccl();
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->comm_stream) {
    cudaStreamSynchronize(h->comm_stream);
    cudaEventDestroy(h->comm_ready);
    cudaEventDestroy(h->comm_done);
    cudaStreamDestroy(h->comm_stream);
    h->comm_stream = nullptr;
    h->comm_ready = h->comm_done = nullptr;
    h->comm_pending = false;
  }
  if (a.ok) a.CommDestroy(static_cast<ncclComm_t>(h->comm));
  h->comm = nullptr;
  h->comm_world = 0;
  return FOP_OK;
}

}  // extern "C"
