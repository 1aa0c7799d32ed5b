// dist.cu -- NCCL plumbing of row-sharded handles (see dist.h).
// NCCL is resolved with dlopen("libnccl.so.2") on first use instead of at link time: when the host process is a
// torch.distributed rank, torch has already loaded ITS bundled libnccl (2.28.x) under that soname and the engine must
// share it -- linking the system copy (2.27.3) into the process first breaks torch's own symbol resolution.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
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
  // optional (pipelined exchange): a second communicator + point-to-point stages
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)(void) = nullptr;
  ncclResult_t (*GroupEnd)(void) = nullptr;
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
  *(void**)(&g_nccl.CommSplit) = dlsym(g_nccl.lib, "ncclCommSplit");
  *(void**)(&g_nccl.Send) = dlsym(g_nccl.lib, "ncclSend");
  *(void**)(&g_nccl.Recv) = dlsym(g_nccl.lib, "ncclRecv");
  *(void**)(&g_nccl.GroupStart) = dlsym(g_nccl.lib, "ncclGroupStart");
  *(void**)(&g_nccl.GroupEnd) = dlsym(g_nccl.lib, "ncclGroupEnd");
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
  // Pipelined exchange (see msdp_dist_exchange_begin): its point-to-point stages run on their own stream and their own
  // communicator, so they can overlap the scalar all-reduces and the product passes of the main stream.
  // Measured on B200 x2 / x8 (profiles/r1_pipelined_exchange.txt), product + exchange per Hessian product:
  //   plain ncclAllGather then one SpMM            2.40 ms (N = 2)   1.25 ms (N = 8)
  //   mode 1: NCCL send/recv stages + passes       2.58 ms           5.98 ms   (p2p stages far below all-gather rate)
  //   mode 2: peer-memory copy stages + passes     1.93 ms           0.83 ms   <- default
  const char* ep = getenv("MANISDP_PIPELINE");
  const int want = ep ? atoi(ep) : 2;
  if (want && g_nccl.CommSplit && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd) {
    ncclComm_t comm2 = nullptr;
    if (g_nccl.CommSplit(comm, 0, h->rank, &comm2, nullptr) == ncclSuccess && comm2) {
      h->nccl_comm2 = (void*)comm2;
      CUDA_TRY(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
      h->ev_stage.resize((size_t)h->world);
      for (auto& e : h->ev_stage) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->pipeline = want >= 2 ? 2 : 1;  // 2: stages are peer-memory copies over CUDA IPC mappings (below)
      if (h->pipeline == 2) {
        CUDA_TRY(h, cudaMalloc((void**)&h->ipc_dev, (size_t)h->world * 256 + 8));  // IPC_SLOT bytes per rank + barrier
        CUDA_TRY(h, cudaMemset(h->ipc_dev, 0, (size_t)h->world * 256 + 8));
        h->peer_d.assign((size_t)h->world, nullptr);
        h->peer_u.assign((size_t)h->world, nullptr);
        h->peer_y[0].assign((size_t)h->world, nullptr);
        h->peer_y[1].assign((size_t)h->world, nullptr);
      }
    }
  }
  return MANISDP_OK;
}

// Arrays every rank publishes to its peers (CUDA IPC): the tCG direction, SLOT_U and the two point buffers.
static const int IPC_K = 4;
static const size_t IPC_SLOT = 64 * IPC_K;  // bytes of handles per rank
static double* ipc_local(manisdp_handle* h, int k) { return k == 0 ? h->d : k == 1 ? h->Uslot : h->Ybuf[k - 2]; }
static std::vector<double*>& ipc_map(manisdp_handle* h, int k) {
  return k == 0 ? h->peer_d : k == 1 ? h->peer_u : h->peer_y[k - 2];
}

static void ipc_close(manisdp_handle* h) {
  for (int k = 0; k < IPC_K; ++k)
    for (auto& p : ipc_map(h, k))
      if (p) cudaIpcCloseMemHandle(p), p = nullptr;
}

// Collective: drop every mapping of peer memory and wait until all ranks have done so.  Must run BEFORE an exported
// array is freed (resize with reallocation, destroy): freeing memory that a peer still has mapped is undefined.
int msdp_dist_ipc_release(manisdp_handle* h) {
  if (h->world <= 1 || h->pipeline != 2 || !h->ipc_dev || !h->nccl_comm) return MANISDP_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->comm_stream) CUDA_TRY(h, cudaStreamSynchronize(h->comm_stream));
  ipc_close(h);
  h->ipc_ready = 0;
  MSDP_TRY(msdp_dist_barrier(h));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MANISDP_OK;
}

// (Re)publish the exchange sources after the work arrays were (re)allocated: every rank exports IPC handles of its
// direction array `d`, of SLOT_U and of the two point buffers, the handles travel through one NCCL all-gather
// (IPC_SLOT bytes per rank), and every rank maps its peers' arrays.  Collective: all ranks resize in lock-step (the
// drivers change p on all ranks alike).
int msdp_dist_ipc_refresh(manisdp_handle* h) {
  if (h->world <= 1 || h->pipeline != 2) return MANISDP_OK;
  h->ipc_ready = 0;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->comm_stream) CUDA_TRY(h, cudaStreamSynchronize(h->comm_stream));
  ipc_close(h);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  unsigned char mine[IPC_SLOT];
  memset(mine, 0, sizeof(mine));
  // a failure to export / map is not an error of the solve: every rank reports it and all fall back to the plain
  // all-gather together (the decision must be collective, the two paths issue different NCCL calls)
  double ok_local = 1.0;
  for (int k = 0; k < IPC_K; ++k) {
    cudaIpcMemHandle_t hk;
    if (cudaIpcGetMemHandle(&hk, ipc_local(h, k)) != cudaSuccess) {
      ok_local = 0.0;
      cudaGetLastError();
      break;
    }
    memcpy(mine + 64 * k, &hk, 64);
  }
  unsigned char* dev = (unsigned char*)h->ipc_dev;
  CUDA_TRY(h, cudaMemcpyAsync(dev + (size_t)h->rank * IPC_SLOT, mine, IPC_SLOT, cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(h, g_nccl.AllGather(dev + (size_t)h->rank * IPC_SLOT, dev, IPC_SLOT, ncclChar, (ncclComm_t)h->nccl_comm,
                               h->stream));
  std::vector<unsigned char> all((size_t)h->world * IPC_SLOT);
  CUDA_TRY(h, cudaMemcpyAsync(all.data(), dev, all.size(), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int q = 0; q < h->world && ok_local != 0.0; ++q) {
    if (q == h->rank) continue;
    for (int k = 0; k < IPC_K; ++k) {
      cudaIpcMemHandle_t hk;
      memcpy(&hk, &all[(size_t)q * IPC_SLOT + 64 * k], 64);
      if (cudaIpcOpenMemHandle((void**)&ipc_map(h, k)[(size_t)q], hk, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok_local = 0.0;
        cudaGetLastError();
        break;
      }
    }
  }
  double* bar = (double*)(dev + (size_t)h->world * IPC_SLOT);
  CUDA_TRY(h, cudaMemcpyAsync(bar, &ok_local, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(h, g_nccl.AllReduce(bar, bar, 1, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  double ok_all = 0.0;
  CUDA_TRY(h, cudaMemcpyAsync(&ok_all, bar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemsetAsync(bar, 0, sizeof(double), h->stream));
  h->ipc_ready = (ok_all > (double)h->world - 0.5) ? 1 : 0;
  if (!h->ipc_ready) ipc_close(h);
  if (h->ipc_ready) {  // device tables for the direct peer-gather kernel: table k at [k*G, (k+1)*G)
    const int G = h->world;
    std::vector<double*> tab((size_t)IPC_K * G);
    for (int k = 0; k < IPC_K; ++k)
      for (int q = 0; q < G; ++q) tab[(size_t)k * G + q] = (q == h->rank) ? ipc_local(h, k) : ipc_map(h, k)[(size_t)q];
    if (!h->peer_tab_dev) CUDA_TRY(h, cudaMalloc((void**)&h->peer_tab_dev, tab.size() * sizeof(double*)));
    CUDA_TRY(h, cudaMemcpy(h->peer_tab_dev, tab.data(), tab.size() * sizeof(double*), cudaMemcpyHostToDevice));
  }
  return MANISDP_OK;
}

const double* const* msdp_dist_peer_table(manisdp_handle* h, const double* local) {
  if (h->pipeline != 2 || !h->ipc_ready || !h->peer_tab_dev) return nullptr;
  for (int k = 0; k < IPC_K; ++k)
    if (local == ipc_local(h, k)) return (const double* const*)(h->peer_tab_dev + (size_t)k * h->world);
  return nullptr;
}

int msdp_dist_barrier(manisdp_handle* h) {
  if (h->world <= 1 || !h->ipc_dev) return MANISDP_OK;
  double* bar = (double*)((unsigned char*)h->ipc_dev + (size_t)h->world * IPC_SLOT);
  NCCL_TRY(h, g_nccl.AllReduce(bar, bar, 1, ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  return MANISDP_OK;
}

void msdp_dist_destroy(manisdp_handle* h) {
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  msdp_dist_ipc_release(h);  // every rank unmaps before any rank frees (the arrays are freed after this returns)
  ipc_close(h);
  if (h->ipc_dev) cudaFree(h->ipc_dev);
  h->ipc_dev = nullptr;
  if (h->peer_tab_dev) cudaFree(h->peer_tab_dev);
  h->peer_tab_dev = nullptr;
  if (h->nccl_comm2 && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm2);
  if (h->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
  h->nccl_comm = h->nccl_comm2 = nullptr;
  for (auto& e : h->ev_stage)
    if (e) cudaEventDestroy(e);
  h->ev_stage.clear();
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  h->ev_ready = nullptr;
  h->comm_stream = nullptr;
  h->pipeline = 0;
}

// Staged all-gather of the thin factor, overlappable with the product: stage 0 copies the own chunk into place, stage s
// (1 <= s < G) sends the own chunk to rank (r + s) mod G and receives the chunk of rank (r - s) mod G, every link busy
// in every stage.  ev_stage[s] fires when the chunk of rank (r - s) mod G has landed in `full`; the product then runs
// as G column passes in that order (spmm.cu), each waiting only for its own chunk.
int msdp_dist_exchange_begin(manisdp_handle* h, const double* local, double* full) {
  const int G = h->world, r = h->rank;
  const size_t cnt = (size_t)(msdp_rows_per_rank(h->n, G) * h->ld);
  ncclComm_t comm2 = (ncclComm_t)h->nccl_comm2;
  // peer-copy stages: rank r PULLS the chunk of rank (r - s) mod G straight out of that rank's memory (NVLink, copy
  // engine).  The pull is not ordered by the owner's stream, so a one-double all-reduce on the main stream first acts
  // as the "every rank has finished writing its operand" barrier; the owner cannot overwrite the operand before all
  // pulls are done because its next writer comes after the product's own scalar all-reduce, which every rank enters
  // only after its last pass, i.e. after its last pull.
  const std::vector<double*>* peers = nullptr;
  if (h->pipeline == 2) {
    if (local == h->d) peers = &h->peer_d;
    if (local == h->Uslot) peers = &h->peer_u;
    if (peers) {
      bool mapped = true;
      for (int q = 0; q < G; ++q) mapped = mapped && (q == r || (*peers)[(size_t)q] != nullptr);
      if (!mapped) peers = nullptr;
    }
  }
  if (peers) {
    MSDP_TRY(msdp_dist_barrier(h));  // every rank has written its chunk
    CUDA_TRY(h, cudaEventRecord(h->ev_ready, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    CUDA_TRY(h, cudaMemcpyAsync(full + (size_t)r * cnt, local, cnt * sizeof(double), cudaMemcpyDeviceToDevice,
                                h->comm_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_stage[0], h->comm_stream));
    for (int s = 1; s < G; ++s) {
      const int src = (r - s + G) % G;
      CUDA_TRY(h, cudaMemcpyAsync(full + (size_t)src * cnt, (*peers)[(size_t)src], cnt * sizeof(double),
                                  cudaMemcpyDefault, h->comm_stream));
      CUDA_TRY(h, cudaEventRecord(h->ev_stage[(size_t)s], h->comm_stream));
    }
    return MANISDP_OK;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev_ready, h->stream));  // `local` is final, previous readers of `full` are done
  CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
  CUDA_TRY(h, cudaMemcpyAsync(full + (size_t)r * cnt, local, cnt * sizeof(double), cudaMemcpyDeviceToDevice,
                              h->comm_stream));
  CUDA_TRY(h, cudaEventRecord(h->ev_stage[0], h->comm_stream));
  for (int s = 1; s < G; ++s) {
    const int dst = (r + s) % G, src = (r - s + G) % G;
    NCCL_TRY(h, g_nccl.GroupStart());
    NCCL_TRY(h, g_nccl.Send(local, cnt, ncclDouble, dst, comm2, h->comm_stream));
    NCCL_TRY(h, g_nccl.Recv(full + (size_t)src * cnt, cnt, ncclDouble, src, comm2, h->comm_stream));
    NCCL_TRY(h, g_nccl.GroupEnd());
    CUDA_TRY(h, cudaEventRecord(h->ev_stage[(size_t)s], h->comm_stream));
  }
  return MANISDP_OK;
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
