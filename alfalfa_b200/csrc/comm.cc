// comm.cc -- the one exchange step of the path (BASELINE.json configs[3]): a GOP that continues on
// another GPU needs the producing GPU's reference rasters (References, decoder.hh:123-141).  The rasters go
// device to device with ncclBroadcast over NVLink, queued on the engine's lane stream between the kernels that
// produced them and the kernels that will read them -- no host synchronisation, no staging copy.  The small
// host-side part of a Decoder (DecoderState, a kilobyte) travels through vp8gpu_comm_broadcast_bytes or any
// channel the caller already has.
//
// NCCL is resolved at run time (dlopen): libvp8gpu.so carries no NCCL dependency, loads on boxes without it, and
// in a process that already holds a copy (torch bundles its own libnccl.so.2) it uses that one.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>

#include <mutex>
#include <string>

#include "../../include/vp8gpu.h"
#include "engine.hpp"

using vp8::Engine;

extern "C" Engine* vp8gpu_ctx_engine(vp8gpu_ctx* ctx);  // capi.cc

namespace {
struct NcclId {
  char internal[128];  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES)
};
struct Nccl {
  void* so = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string why;
};
Nccl* nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    n.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // a copy the process already loaded (torch's)
    if (!n.so) n.so = dlopen("libnccl.so.2", RTLD_NOW);
    if (!n.so) n.so = dlopen("libnccl.so", RTLD_NOW);
    if (!n.so) {
      n.why = "libnccl.so.2 not found";
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(n.so, name);
      if (!p && n.why.empty()) n.why = std::string("libnccl lacks ") + name;
      return p;
    };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.Broadcast = reinterpret_cast<decltype(n.Broadcast)>(sym("ncclBroadcast"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &n;
}
constexpr int kNcclUint8 = 1;  // ncclDataType_t (nccl.h)
}  // namespace

struct vp8gpu_comm {
  vp8gpu_ctx* ctx = nullptr;
  Engine* e = nullptr;
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  uint8_t* d_bytes = nullptr;  // staging for vp8gpu_comm_broadcast_bytes
  size_t d_cap = 0;
};

extern "C" {

int vp8gpu_comm_unique_id(uint8_t out[128]) {
  Nccl* n = nccl();
  if (!out) return VP8GPU_ERR_LOGIC;
  if (!n->GetUniqueId) return VP8GPU_ERR_UNSUPPORTED;
  NcclId id;
  if (n->GetUniqueId(&id) != 0) return VP8GPU_ERR_CUDA;
  memcpy(out, id.internal, 128);
  return VP8GPU_OK;
}

int vp8gpu_comm_create(vp8gpu_ctx* ctx, int rank, int nranks, const uint8_t unique_id[128], vp8gpu_comm** out) {
  if (!ctx || !unique_id || !out || rank < 0 || rank >= nranks) return VP8GPU_ERR_LOGIC;
  Engine* e = vp8gpu_ctx_engine(ctx);
  Nccl* n = nccl();
  if (!n->CommInitRank) return e->fail(VP8GPU_ERR_UNSUPPORTED, "NCCL unavailable: " + n->why);
  cudaSetDevice(e->device());
  NcclId id;
  memcpy(id.internal, unique_id, 128);
  void* comm = nullptr;
  const int rc = n->CommInitRank(&comm, nranks, id, rank);
  if (rc != 0) return e->fail(VP8GPU_ERR_CUDA, std::string("ncclCommInitRank: ") + n->GetErrorString(rc));
  vp8gpu_comm* c = new vp8gpu_comm();
  c->ctx = ctx, c->e = e, c->comm = comm, c->rank = rank, c->nranks = nranks;
  *out = c;
  return VP8GPU_OK;
}

void vp8gpu_comm_destroy(vp8gpu_comm* c) {
  if (!c) return;
  cudaSetDevice(c->e->device());
  c->e->sync_all();
  if (c->d_bytes) cudaFree(c->d_bytes);
  if (c->comm) nccl()->CommDestroy(c->comm);
  delete c;
}

int vp8gpu_comm_broadcast_frames(vp8gpu_comm* c, int root, int lane, const vp8gpu_frame_id* ids, int n) {
  if (!c || !ids || n < 0 || n > 32 || root < 0 || root >= c->nranks) return VP8GPU_ERR_LOGIC;
  if (n == 0) return VP8GPU_OK;
  Engine* e = c->e;
  Nccl* nc = nccl();
  cudaSetDevice(e->device());
  // stream order: the root's rasters are read after the kernels that wrote them, the receivers' rasters are
  // written after every earlier user; later decodes on any lane wait for this lane through the raster events
  const uint32_t write_mask = c->rank == root ? 0u : ~0u;
  int rc = e->acquire_frames(lane, ids, n, write_mask);
  if (rc != VP8GPU_OK) return rc;
  cudaStream_t s = e->stream(lane);
  const size_t bytes = e->geom().frame_bytes;
  int nrc = nc->GroupStart();
  for (int i = 0; i < n && nrc == 0; i++) {
    uint8_t* p = e->frame_dev(ids[i]);
    nrc = nc->Broadcast(p, p, bytes, kNcclUint8, root, c->comm, s);
  }
  const int end_rc = nc->GroupEnd();
  if (nrc == 0) nrc = end_rc;
  if (nrc != 0) return e->fail(VP8GPU_ERR_CUDA, std::string("ncclBroadcast: ") + nc->GetErrorString(nrc));
  e->count_launches(n);
  return e->mark_frames(lane, ids, n, write_mask);
}

int vp8gpu_comm_broadcast_bytes(vp8gpu_comm* c, int root, void* buf, size_t bytes) {
  if (!c || (!buf && bytes) || root < 0 || root >= c->nranks) return VP8GPU_ERR_LOGIC;
  if (bytes == 0) return VP8GPU_OK;
  Engine* e = c->e;
  Nccl* nc = nccl();
  cudaSetDevice(e->device());
  if (int rc = e->ensure_lane(0)) return rc;
  if (c->d_cap < bytes) {
    if (c->d_bytes) cudaFree(c->d_bytes);
    c->d_cap = bytes + bytes / 2 + 4096;
    if (cudaMalloc(&c->d_bytes, c->d_cap) != cudaSuccess) {
      c->d_bytes = nullptr, c->d_cap = 0;
      return e->fail(VP8GPU_ERR_NOMEM, "comm staging allocation failed");
    }
  }
  cudaStream_t s = e->stream(0);
  if (c->rank == root && cudaMemcpyAsync(c->d_bytes, buf, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess)
    return e->fail(VP8GPU_ERR_CUDA, "comm staging upload failed");
  const int nrc = nc->Broadcast(c->d_bytes, c->d_bytes, bytes, kNcclUint8, root, c->comm, s);
  if (nrc != 0) return e->fail(VP8GPU_ERR_CUDA, std::string("ncclBroadcast: ") + nc->GetErrorString(nrc));
  if (c->rank != root && cudaMemcpyAsync(buf, c->d_bytes, bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return e->fail(VP8GPU_ERR_CUDA, "comm staging download failed");
  if (cudaStreamSynchronize(s) != cudaSuccess) return e->fail(VP8GPU_ERR_CUDA, "comm stream failed");
  return VP8GPU_OK;
}

int vp8gpu_comm_rank(const vp8gpu_comm* c) { return c ? c->rank : -1; }
int vp8gpu_comm_size(const vp8gpu_comm* c) { return c ? c->nranks : 0; }

}  // extern "C"
