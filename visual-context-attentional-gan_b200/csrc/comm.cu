// NCCL helpers of the C ABI: the data-parallel gradient exchange of train.py:112-119's replacement -- a sum
// all-reduce of flat gradient buckets over NVLink -- callable without torch.distributed.  NCCL is bound at run time
// (dlopen of the libnccl.so.2 already in the process, torch's bundled one when torch is loaded, else the system's): the
// library carries no link-time dependency on it, and single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);

struct NcclApi {
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
      api.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
      api.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
      api.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
      api.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
      api.ok = api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy;
    }
  }
  return api;
}

constexpr int MAX_DEV = 64;
NcclComm g_comm[MAX_DEV] = {};
int g_world[MAX_DEV] = {};

int fail(const char* what, int rc) {
  NcclApi& a = nccl();
  vca_set_error("%s failed: %s", what, a.error_string ? a.error_string(rc) : "NCCL error");
  return VCA_ERR_CUDA;
}

}  // namespace

extern "C" {

// 1 when an NCCL library could be bound in this process.
int vca_comm_available() { return nccl().ok ? 1 : 0; }

// Rank 0 creates the 128-byte rendezvous token; the caller ships it to the other ranks (any side channel: a file, MPI,
// torch.distributed's store) and every rank then calls vca_comm_init with it.
int vca_comm_unique_id(void* id128) {
  VCA_CHECK_ARG(id128);
  if (!nccl().ok) { vca_set_error("NCCL is not available in this process (libnccl.so.2 could not be loaded)"); return VCA_ERR_UNSUPPORTED; }
  const int rc = nccl().get_unique_id(reinterpret_cast<NcclUniqueId*>(id128));
  return rc ? fail("ncclGetUniqueId", rc) : VCA_OK;
}

// One communicator per device of this process (one process per GPU): bound to the CURRENT device.  Collective: every rank
// of the job must call it with the same token and world size.
int vca_comm_init(const void* unique_id, int rank, int world) {
  VCA_CHECK_ARG(unique_id && world > 0 && rank >= 0 && rank < world);
  if (!nccl().ok) { vca_set_error("NCCL is not available in this process (libnccl.so.2 could not be loaded)"); return VCA_ERR_UNSUPPORTED; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) { vca_set_error("vca_comm_init: no current CUDA device"); return VCA_ERR_CUDA; }
  if (g_comm[dev]) { nccl().comm_destroy(g_comm[dev]); g_comm[dev] = nullptr; }
  NcclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  const int rc = nccl().comm_init_rank(&g_comm[dev], world, id, rank);
  if (rc) { g_comm[dev] = nullptr; return fail("ncclCommInitRank", rc); }
  g_world[dev] = world;
  return VCA_OK;
}

// In-place SUM all-reduce of `count` elements (dtype VCA_F32 or VCA_BF16) over the current device's communicator, enqueued
// on `stream` (no host synchronisation).  The 1/world of the gradient mean is folded into vca_adam_step's grad_scale.
int vca_allreduce_bucket(void* ptr, long long count, int dtype, cudaStream_t stream) {
  VCA_CHECK_ARG(ptr && count > 0 && (dtype == VCA_F32 || dtype == VCA_BF16));
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV || !g_comm[dev]) {
    vca_set_error("vca_allreduce_bucket: vca_comm_init has not been called on this device"); return VCA_ERR_ARG;
  }
  const int nccl_dtype = dtype == VCA_F32 ? 7 /* ncclFloat32 */ : 9 /* ncclBfloat16 */;
  const int rc = nccl().all_reduce(ptr, ptr, (size_t)count, nccl_dtype, 0 /* ncclSum */, g_comm[dev], stream);
  return rc ? fail("ncclAllReduce", rc) : VCA_OK;
}

// World size of the current device's communicator (0: none).
int vca_comm_world() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 0;
  return g_comm[dev] ? g_world[dev] : 0;
}

int vca_comm_destroy() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return VCA_OK;
  if (g_comm[dev]) { nccl().comm_destroy(g_comm[dev]); g_comm[dev] = nullptr; g_world[dev] = 0; }
  return VCA_OK;
}

}  // extern "C"
