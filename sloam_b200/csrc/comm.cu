// comm.cu -- multi-GPU: the one exchange of the sharded path.
//
// Keyframes are independent once firstScan_ / prevGPlanes_ / the submap are explicit inputs
// (SURVEY 8(e)), so every rank (one process per GPU) runs the fused path on its own block of
// keyframes and nothing is exchanged during compute.  What the caller of a sharded sequence
// needs back is the per-keyframe result of every rank: SloamOutput (sloam/include/core/sloam.h:
// 48-55) -- pose, T_Delta, status, counts, matches[], tm[] -- gathered here with ncclAllGather
// over NVLink on a SIDE stream: the gather of batch i waits for the kernels of batch i, the
// kernels of batch i + 1 start at once and only its last kernels (which overwrite the output
// buffers the gather reads) wait for the gather to finish.
//
// NCCL is opened at run time (dlopen libnccl.so.2: in a torch process that is the library
// torch already loaded), so the CUDA library itself has no link-time dependency on it and
// loads on a box without NCCL; only the five entry points below are used.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace sb {

// the parts of nccl.h this file uses (stable since NCCL 2.0)
typedef struct ncclComm *nccl_comm_t;
struct nccl_unique_id { char internal[128]; };
static_assert(sizeof(nccl_unique_id) == SLOAM_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
constexpr int kNcclUint8 = 1;  // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(nccl_unique_id *) = nullptr;
  int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok() const { return lib && GetUniqueId && CommInitRank && CommDestroy && AllGather && GroupStart && GroupEnd; }
};

static NcclApi &nccl() {
  static NcclApi api;
  if (!api.lib) {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.lib, "ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.lib, "ncclGroupEnd"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    }
  }
  return api;
}

struct CommState {
  nccl_comm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;  // the side stream the gathers run on
  cudaEvent_t ev_ready = nullptr; // batch kernels done -> gather may start
};

static int nccl_err(sloam_ctx *c, int rc, const char *what) {
  NcclApi &a = nccl();
  return set_err(c, SLOAM_E_CUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
}

}  // namespace sb

using namespace sb;

extern "C" {

int sloam_b200_comm_unique_id(void *id128) {
  if (!id128) return SLOAM_E_INVALID;
  NcclApi &a = nccl();
  if (!a.ok()) return SLOAM_E_NODEVICE;  // no NCCL on this box
  nccl_unique_id id;
  if (a.GetUniqueId(&id) != 0) return SLOAM_E_CUDA;
  std::memcpy(id128, &id, sizeof id);
  return SLOAM_OK;
}

int sloam_b200_comm_init(sloam_ctx *c, int rank, int world, const void *id128) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return set_err(c, SLOAM_E_INVALID, "comm_init: bad arguments");
  NcclApi &a = nccl();
  if (!a.ok()) return set_err(c, SLOAM_E_NODEVICE, "comm_init: libnccl.so.2 not found");
  if (c->comm) sloam_b200_comm_destroy(c);
  cudaSetDevice(c->device);
  CommState *s = new (std::nothrow) CommState();
  if (!s) return SLOAM_E_NOMEM;
  s->rank = rank;
  s->world = world;
  nccl_unique_id id;
  std::memcpy(&id, id128, sizeof id);
  const int rc = a.CommInitRank(&s->comm, world, id, rank);
  if (rc != 0) { delete s; return nccl_err(c, rc, "ncclCommInitRank"); }
  // highest priority: when a gather and the first kernels of the next batch become runnable
  // together, the few CTAs of the gather are placed first instead of behind a full grid
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      (!c->ev_gather_done && cudaEventCreateWithFlags(&c->ev_gather_done, cudaEventDisableTiming) != cudaSuccess)) {
    a.CommDestroy(s->comm);
    delete s;
    return set_err(c, SLOAM_E_CUDA, "comm_init: stream / event creation failed");
  }
  c->comm = s;
  return SLOAM_OK;
}

int sloam_b200_comm_destroy(sloam_ctx *c) {
  if (!c || !c->comm) return SLOAM_OK;
  CommState *s = static_cast<CommState *>(c->comm);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(s->stream);
  if (s->comm) nccl().CommDestroy(s->comm);
  if (s->ev_ready) cudaEventDestroy(s->ev_ready);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  c->comm = nullptr;
  c->gather_pending = false;
  return SLOAM_OK;
}

int sloam_b200_comm_size(const sloam_ctx *c) { return (c && c->comm) ? static_cast<const CommState *>(c->comm)->world : 1; }
int sloam_b200_comm_rank(const sloam_ctx *c) { return (c && c->comm) ? static_cast<const CommState *>(c->comm)->rank : 0; }

int sloam_b200_gather_results_dev(sloam_ctx *c, int K, const sloam_batch_out *local, const sloam_batch_out *all) {
  if (!c || !c->comm || K <= 0 || !local || !all || !local->results || !all->results)
    return set_err(c, SLOAM_E_INVALID, "gather_results: bad arguments / comm_init not called");
  CommState *s = static_cast<CommState *>(c->comm);
  NcclApi &a = nccl();
  const size_t T = (size_t)c->hp.p.max_trees, k = (size_t)K;
  // the gather reads what the kernels queued so far on the context stream have written
  SB_CUDA(c, cudaEventRecord(s->ev_ready, c->stream));
  SB_CUDA(c, cudaStreamWaitEvent(s->stream, s->ev_ready, 0));
  int rc = a.GroupStart();
  if (rc != 0) return nccl_err(c, rc, "ncclGroupStart");
  struct Part { const void *src; void *dst; size_t bytes; };
  const Part parts[4] = {
      {local->results, all->results, k * sizeof(sloam_kf_result)},
      {local->matches, all->matches, k * T * sizeof(int32_t)},
      {local->tm, all->tm, k * T * sizeof(sloam_cylinder)},
      {local->tm_id, all->tm_id, k * T * sizeof(int32_t)}};
  for (const Part &p : parts) {
    if (!p.src || !p.dst) continue;  // optional arrays
    rc = a.AllGather(p.src, p.dst, p.bytes, kNcclUint8, s->comm, s->stream);
    if (rc != 0) { a.GroupEnd(); return nccl_err(c, rc, "ncclAllGather"); }
  }
  rc = a.GroupEnd();
  if (rc != 0) return nccl_err(c, rc, "ncclGroupEnd");
  SB_CUDA(c, cudaEventRecord(c->ev_gather_done, s->stream));
  c->gather_pending = true;  // the next fused run keeps its output kernels behind this event (pipeline.cu)
  return SLOAM_OK;
}

int sloam_b200_comm_wait(sloam_ctx *c) {
  if (!c) return SLOAM_E_INVALID;
  if (c->gather_pending) {
    SB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_gather_done, 0));
    c->gather_pending = false;
  }
  return SLOAM_OK;
}

}  // extern "C"
