// dist.cu -- NCCL plumbing of row-sharded handles (see dist.h).
// NCCL is resolved with dlopen("libnccl.so.2") on first use instead of at link time: when the host process is a
// torch.distributed rank, torch has already loaded ITS bundled libnccl (2.28.x) under that soname and the engine must
// share it -- linking the system copy (2.27.3) into the process first breaks torch's own symbol resolution.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include "dist.h"
#include "kernels.cuh"

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

bool nccl_load(std::string& why) {
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return false;
  }
#define SYM(field, name)                                               \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                  \
  if (!g_nccl.field) {                                                 \
    why = std::string("libnccl lacks ") + name;                        \
    return false;                                                      \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.ok = true;
  return true;
}
}  // namespace

#define NCCL_TRY(h, expr)                                                                                  \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess)                                                                                 \
      return msdp_fail(h, MANISDP_E_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));          \
  } while (0)

int msdp_dist_init(manisdp_handle* h, const void* unique_id) {
  if (h->world <= 1) return MANISDP_OK;
  if (!unique_id) return msdp_fail(h, MANISDP_E_ARG, "sharded handle needs nccl_unique_id");
  std::string why;
  if (!nccl_load(why)) return msdp_fail(h, MANISDP_E_NCCL, why);
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  NCCL_TRY(h, g_nccl.CommInitRank(&comm, h->world, id, h->rank));
  h->nccl_comm = (void*)comm;
  return MANISDP_OK;
}

void msdp_dist_destroy(manisdp_handle* h) {
  if (h->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
  h->nccl_comm = nullptr;
}

int msdp_dist_allgather_rows(manisdp_handle* h, const double* local, double* full) {
  const int64_t rpr = msdp_rows_per_rank(h->n, h->world);
  NCCL_TRY(h, g_nccl.AllGather(local, full, (size_t)(rpr * h->ld), ncclDouble, (ncclComm_t)h->nccl_comm, h->stream));
  return MANISDP_OK;
}

int msdp_dist_allgather_block(manisdp_handle* h, const double* src, double* dst, int64_t count) {
  NCCL_TRY(h, g_nccl.AllGather(src, dst, (size_t)count, ncclDouble, (ncclComm_t)h->nccl_comm, h->stream));
  return MANISDP_OK;
}

int msdp_dist_allreduce_tmp(manisdp_handle* h, int count) {
  NCCL_TRY(h, g_nccl.AllReduce(h->st->tmp, h->st->tmp, (size_t)count, ncclDouble, ncclSum,
                               (ncclComm_t)h->nccl_comm, h->stream));
  return MANISDP_OK;
}

int msdp_dist_allreduce_buf(manisdp_handle* h, double* buf, int64_t count) {
  if (h->world <= 1) return MANISDP_OK;
  NCCL_TRY(h, g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  return MANISDP_OK;
}

__global__ void k_finish_init(RtrState* st) {
  st->fx = 0.5 * st->tmp[0];
  st->gradnorm2 = st->tmp[1];
}
int msdp_dist_finish_init(manisdp_handle* h) {
  k_finish_init<<<1, 1, 0, h->stream>>>(h->st);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// exported helper so the host language can create the id on rank 0 and broadcast it with its own transport
extern "C" int manisdp_nccl_unique_id(void* out128) {
  std::string why;
  if (!out128 || !nccl_load(why)) return MANISDP_E_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MANISDP_E_NCCL;
  memcpy(out128, &id, sizeof(id));
  return MANISDP_OK;
}
