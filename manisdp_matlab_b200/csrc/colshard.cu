// colshard.cu -- COLUMN-sharded trust-region solve of the ONLYUNITDIAG path (the "p-sharded" layout of SURVEY 8e).
//
// Row sharding (dist.cu) has to move the whole thin factor to every GPU before every Hessian product: at n = 1e6,
// p = 64 that is 448 MB per GPU per product, ~0.6 ms over NVLink 5 against ~0.46 ms of local product time at 8 GPUs --
// the exchange, not the kernel, bounded the 8-GPU efficiency at 0.53 (round 1).  Here every rank keeps ALL n rows of
// C and of Y but only pl = ceil(p / G) COLUMNS of every n x p array:
//   * the product C * U_loc needs no communication at all (the rows of the operand a rank gathers are its own);
//   * everything that couples the columns of a row is a per-row scalar: the oblique projections
//     (ManiSDP_onlyunitdiag.m:124,129,139-141 -- sum(Y.*eH), sum(YC.*Y)) and the row norms of the retraction (:145-148).
//     Each rank forms its partial row sums, one NCCL all-reduce of an n-vector (8 MB at n = 1e6, NVLS-reduced inside the
//     switch) completes them, and a second light pass over the rank's columns applies them;
//   * the tCG inner products are sums over all entries, all-reduced as scalars exactly like in the row-sharded path.
// Per tCG iteration a rank moves 3 n-vectors + 7 scalars through the switch, in three all-reduces (row sums of the
// product; <mdelta, Hmdelta>; the update packet with the two row-sum vectors of the next direction), instead of the
// n x p factor.
//
// A column-sharded handle is an ordinary single-GPU handle (h->world == 1: every rank holds every row) that is SPLIT for
// the trust-region solve and MERGED for the outer-loop steps (KKT / eigen step / rank step / escape), which then run
// redundantly and deterministically on every rank on the full-width factor:
//   manisdp_col_split : keep columns [rank*pl, (rank+1)*pl) of the current point (zero-padded past p)
//   manisdp_col_merge : all-gather the column slices -> full factor of width G*pl
// While split only tr_solve / cost / get-set of the local slice are meaningful.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "dist.h"
#include "kernels.cuh"
#include "rowops.cuh"
#include "scalar_logic.cuh"

// ---- NCCL (same lazy dlopen as dist.cu: the process may already hold torch's bundled libnccl) -------------------------
namespace {
struct ColNccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
} g_cn;

std::mutex g_cn_mu;  // (the workers of a single-process group create their handles concurrently)
bool col_nccl_load(std::string& why) {
  std::lock_guard<std::mutex> lk(g_cn_mu);
  if (g_cn.ok) return true;
  for (const char* nm : {"libnccl.so.2", "libnccl.so"}) {
    g_cn.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_cn.lib) break;
  }
  if (!g_cn.lib) {
    why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return false;
  }
  *(void**)(&g_cn.CommInitRank) = dlsym(g_cn.lib, "ncclCommInitRank");
  *(void**)(&g_cn.CommDestroy) = dlsym(g_cn.lib, "ncclCommDestroy");
  *(void**)(&g_cn.AllGather) = dlsym(g_cn.lib, "ncclAllGather");
  *(void**)(&g_cn.AllReduce) = dlsym(g_cn.lib, "ncclAllReduce");
  *(void**)(&g_cn.GetErrorString) = dlsym(g_cn.lib, "ncclGetErrorString");
  if (!g_cn.CommInitRank || !g_cn.CommDestroy || !g_cn.AllGather || !g_cn.AllReduce || !g_cn.GetErrorString) {
    why = "libnccl lacks a required symbol";
    return false;
  }
  g_cn.ok = true;
  return true;
}
}  // namespace

#define CNCCL_TRY(h, expr)                                                                                 \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess)                                                                                 \
      return msdp_fail(h, MANISDP_E_NCCL, std::string(#expr) + ": " + g_cn.GetErrorString(_r));            \
  } while (0)

int msdp_col_init(manisdp_handle* h, const void* unique_id, int world, int rank) {
  h->cworld = world > 1 ? world : 1;
  h->crank = world > 1 ? rank : 0;
  h->col_mode = 1;
  if (const char* e = getenv("MANISDP_COL_GRAPH")) h->col_graph = atoi(e);
  CUDA_TRY(h, cudaMalloc((void**)&h->col_rowvec, (size_t)h->n * sizeof(double)));
  CUDA_TRY(h, cudaMemset(h->col_rowvec, 0, (size_t)h->n * sizeof(double)));
  CUDA_TRY(h, cudaMalloc((void**)&h->col_pack, (size_t)(8 + 2 * h->n) * sizeof(double)));
  CUDA_TRY(h, cudaMemset(h->col_pack, 0, (size_t)(8 + 2 * h->n) * sizeof(double)));
  if (h->cworld <= 1) return MANISDP_OK;
  if (!unique_id) return msdp_fail(h, MANISDP_E_ARG, "column-sharded handle needs nccl_unique_id");
  std::string why;
  if (!col_nccl_load(why)) return msdp_fail(h, MANISDP_E_NCCL, why);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  CNCCL_TRY(h, g_cn.CommInitRank(&comm, h->cworld, id, h->crank));
  h->col_comm = (void*)comm;
  return MANISDP_OK;
}

void msdp_col_destroy(manisdp_handle* h) {
  if (h->col_comm && g_cn.ok) g_cn.CommDestroy((ncclComm_t)h->col_comm);
  h->col_comm = nullptr;
  if (h->col_rowvec) cudaFree(h->col_rowvec);
  h->col_rowvec = nullptr;
  if (h->col_pack) cudaFree(h->col_pack);
  h->col_pack = nullptr;
}

static int col_allreduce(manisdp_handle* h, double* buf, int64_t count) {
  if (h->cworld <= 1) return MANISDP_OK;
  CNCCL_TRY(h, g_cn.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)h->col_comm, h->stream));
  return MANISDP_OK;
}

// ---- kernels: one GS-lane group per row of the local n x ld_loc slice ----------------------------------------------------
// which point buffer: sel 0 -> explicit pointer, 1 -> current point st->pt (device-selected), 2 -> proposal st->pt ^ 1
__device__ __forceinline__ const double* col_point(const VecPtrs& v, const RtrState* st, const double* expl, int sel) {
  if (sel == 0) return expl;
  const int w = (sel == 1) ? st->pt : (st->pt ^ 1);
  return w ? v.Y1 : v.Y0;
}

// t[row] = sum over the local columns of A[row,:] .* B[row,:]   (partial row sums of sum(Y.*eH), sum(YC.*Y))
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_rowdot(const double* Aexpl, int selA, const double* __restrict__ B, double* __restrict__ t, VecPtrs v,
                 RtrState* st, int64_t nrows, int ld, int in_tcg) {
  if (in_tcg && st->stop != 0) return;
  const double* __restrict__ A = col_point(v, st, Aexpl, selA);
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const double2 a = ld2(A + base + 2 * c), b = ld2(B + base + 2 * c);
        dot += a.x * b.x + a.y * b.y;
      }
    }
    dot = group_sum<GS>(dot, mask);
    if (gl == 0) t[row] = dot;
  }
}

// Hd = raw - Y .* t - D .* eG  (ManiSDP_onlyunitdiag.m:129 with the completed row sums t); q0 = <D, Hd> over the slice
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_hess_finish(const double* Yexpl, const double* eGexpl, const double* __restrict__ D, double* __restrict__ Hd,
                      const double* __restrict__ t, VecPtrs v, RtrState* st, double* partials, int64_t nrows, int ld,
                      int from_state, int tail_mode) {
  __shared__ double sm[32];
  if (tail_mode != TAIL_NONE && st->stop != 0) return;
  const double* __restrict__ Y = col_point(v, st, Yexpl, from_state ? 1 : 0);
  const double* __restrict__ eG = from_state ? (st->pt ? v.eG1 : v.eG0) : eGexpl;
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  double q[1] = {0.0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    const double dot = t[row], eg = eG[row];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const double2 raw = ld2(Hd + base + 2 * c), y = ld2(Y + base + 2 * c), u = ld2(D + base + 2 * c);
        double2 hv;
        hv.x = raw.x - y.x * dot - u.x * eg;
        hv.y = raw.y - y.y * dot - u.y * eg;
        st2(Hd + base + 2 * c, hv);
        q[0] += u.x * hv.x + u.y * hv.y;  // <mdelta, Hmdelta>, tCG.m:166
      }
    }
  }
  if (tail_mode == TAIL_NONE) return;
  double tot[1];
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) st->tmp[0] = tot[0];  // local share; all-reduced, then k_tcg_after_hv_scalar
  }
}

// G = raw - Y .* eG  (:124); tmp[0] = sum(eG) on rank 0 only (every rank holds the complete eG), tmp[1] = |G_loc|^2
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_grad_finish(const double* __restrict__ Y, double* __restrict__ G, const double* __restrict__ eG, RtrState* st,
                      double* partials, int64_t nrows, int ld, int count_cost) {
  __shared__ double sm[2 * 32];
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  double q[2] = {0.0, 0.0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    const double eg = eG[row];
    if (gl == 0 && count_cost) q[0] += eg;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const double2 raw = ld2(G + base + 2 * c), y = ld2(Y + base + 2 * c);
        double2 g;
        g.x = raw.x - y.x * eg;
        g.y = raw.y - y.y * eg;
        st2(G + base + 2 * c, g);
        q[1] += g.x * g.x + g.y * g.y;
      }
    }
  }
  double tot[2];
  if (grid_sum_last<2>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      st->tmp[0] = tot[0];  // un-halved sum(eG) (rank 0) / 0 (others): the scalar kernels halve after the all-reduce
      st->tmp[1] = tot[1];
    }
  }
}

// retraction, first half: Yprop = Y + eta (current point and tCG iterate chosen on the device), t[row] = partial |.|^2
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_retract_raw(VecPtrs v, RtrState* st, double* __restrict__ t, int64_t nrows, int ld) {
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const double* __restrict__ E = st->eta_cur ? v.eta1 : v.eta0;
  double* __restrict__ out = st->pt ? v.Y0 : v.Y1;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double ss = 0.0;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const double2 y = ld2(Y + base + 2 * c), e = ld2(E + base + 2 * c);
        double2 x;
        x.x = y.x + e.x;
        x.y = y.y + e.y;
        st2(out + base + 2 * c, x);
        ss += x.x * x.x + x.y * x.y;
      }
    }
    ss = group_sum<GS>(ss, mask);
    if (gl == 0) t[row] = ss;
  }
}
// second half: Yprop(row,:) /= sqrt(t[row])   (ManiSDP_onlyunitdiag.m:146-147)
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_rowscale(VecPtrs v, RtrState* st, const double* __restrict__ t, int64_t nrows, int ld) {
  double* __restrict__ out = st->pt ? v.Y0 : v.Y1;
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    const double s = 1.0 / sqrt(t[row]);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        double2 x = ld2(out + base + 2 * c);
        x.x *= s;
        x.y *= s;
        st2(out + base + 2 * c, x);
      }
    }
  }
}

// direction, first half: d = r + beta*d (tCG.m:273), t[row] = partial sum(Y .* d)
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_dir_raw(VecPtrs v, RtrState* st, double* __restrict__ t, int64_t nrows, int ld) {
  if (st->stop != 0) return;
  const double beta = st->beta;
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const double2 r = ld2(v.r + base + 2 * c), d = ld2(v.d + base + 2 * c), y = ld2(Y + base + 2 * c);
        double2 dn;
        dn.x = r.x + beta * d.x;
        dn.y = r.y + beta * d.y;
        st2(v.d + base + 2 * c, dn);
        dot += y.x * dn.x + y.y * dn.y;
      }
    }
    dot = group_sum<GS>(dot, mask);
    if (gl == 0) t[row] = dot;
  }
}
// second half: d -= Y .* t  (the tangent re-projection of tCG.m:283, ManiSDP_onlyunitdiag.m:139-141)
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_dir_finish(VecPtrs v, RtrState* st, const double* __restrict__ t, int64_t nrows, int ld) {
  if (st->stop != 0) return;
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    const double dot = t[row];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        double2 d = ld2(v.d + base + 2 * c);
        const double2 y = ld2(Y + base + 2 * c);
        d.x -= y.x * dot;
        d.y -= y.y * dot;
        st2(v.d + base + 2 * c, d);
      }
    }
  }
}

// ---- closures of the split handle ------------------------------------------------------------------------------------
// Hess f(Y)[D] on the local columns; from_state / tail_mode as msdp_maxcut_hess.  D and Hout are n x ld_loc arrays.
int msdp_col_hess(manisdp_handle* h, const double* D, double* Hout, int from_state, int tail_mode) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  // raw product (no communication: every operand row a rank gathers is its own)
  MSDP_TRY(msdp_spmm_shift_tcg(h, D, Hout, ld, tail_mode != TAIL_NONE));
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_rowdot<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(h->Ybuf[h->pt], from_state ? 1 : 0, Hout, h->col_rowvec, v,
                                                               h->st, h->nloc, ld, tail_mode != TAIL_NONE);
  });
  KERNEL_CHECK(h);
  MSDP_TRY(col_allreduce(h, h->col_rowvec, h->nloc));
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_hess_finish<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(h->Ybuf[h->pt], h->eG[h->pt], D, Hout, h->col_rowvec,
                                                                    v, h->st, h->partials, h->nloc, ld, from_state,
                                                                    tail_mode);
  });
  KERNEL_CHECK(h);
  if (tail_mode != TAIL_NONE) {
    MSDP_TRY(col_allreduce(h, h->st->tmp, 1));
    MSDP_TRY(msdp_launch_tcg_after_hv_scalar(h));
  }
  return MANISDP_OK;
}

// cost + gradient at point buffer `buf` (host-known: the host mirror of pt is exact between TR iterations).
// Leaves the all-reduced (sum(eG), |G|^2) in st->tmp[0..1]; the caller finishes like the row-sharded path.
int msdp_col_costgrad(manisdp_handle* h, int buf) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  MSDP_TRY(msdp_spmm_shift_tcg(h, h->Ybuf[buf], h->Gbuf[buf], ld, 0));
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_rowdot<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(h->Ybuf[buf], 0, h->Gbuf[buf], h->eG[buf], v, h->st,
                                                               h->nloc, ld, 0);
  });
  KERNEL_CHECK(h);
  MSDP_TRY(col_allreduce(h, h->eG[buf], h->nloc));  // eG(row) = sum(YC.*Y) over ALL columns, :119
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_grad_finish<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(h->Ybuf[buf], h->Gbuf[buf], h->eG[buf], h->st,
                                                                    h->partials, h->nloc, ld, h->crank == 0 ? 1 : 0);
  });
  KERNEL_CHECK(h);
  return col_allreduce(h, h->st->tmp, 2);
}

int msdp_col_retract(manisdp_handle* h) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_retract_raw<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->col_rowvec, h->nloc, ld);
  });
  KERNEL_CHECK(h);
  MSDP_TRY(col_allreduce(h, h->col_rowvec, h->nloc));
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_rowscale<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->col_rowvec, h->nloc, ld);
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_col_tcg_dir(manisdp_handle* h) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_dir_raw<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->col_rowvec, h->nloc, ld);
  });
  KERNEL_CHECK(h);
  MSDP_TRY(col_allreduce(h, h->col_rowvec, h->nloc));
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_dir_finish<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->col_rowvec, h->nloc, ld);
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}


// ---- fused update pass of the split loop ---------------------------------------------------------------------------------
// The update pass of tcg.cu (eta' = eta - a*mdelta, r' = r - a*Hmdelta and the inner products, tCG.m:192-241) in row
// geometry, which also leaves the partial row sums the NEXT direction needs: mdelta' = P_Y(r' + beta*mdelta) subtracts
// Y .* rowsum(Y .* (r' + beta*mdelta)) = Y .* (t1 + beta*t2) with t1 = rowsum(Y.*r'), t2 = rowsum(Y.*mdelta) -- both are
// known before beta is.  pack = [6 scalars, 2 pad | t1 (n) | t2 (n)] travels through ONE all-reduce per iteration
// (instead of a scalar packet and a separate n-vector for the direction).
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_tcg_update(VecPtrs v, RtrState* st, double* partials, double* __restrict__ pack, int64_t nrows, int ld) {
  __shared__ double sm[MSDP_NQ * 32];
  if (st->stop != 0) return;
  const int branch = st->branch, cur = st->eta_cur;
  const double a = (branch != 0) ? st->tau : st->alpha;
  const double* __restrict__ g = st->pt ? v.G1 : v.G0;
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const double* __restrict__ eo = cur ? v.eta1 : v.eta0;
  double* __restrict__ en = cur ? v.eta0 : v.eta1;
  double* __restrict__ t1 = pack + 8;
  double* __restrict__ t2 = pack + 8 + nrows;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  double q[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const size_t i = base + 2 * c;
        const double2 e = ld2(eo + i), d = ld2(v.d + i), r = ld2(v.r + i), hd = ld2(v.Hd + i), gv = ld2(g + i);
        double2 n;
        n.x = e.x - a * d.x;  // tCG.m:192 / :215
        n.y = e.y - a * d.y;
        st2(en + i, n);
        if (branch != 0) {
          const double hx = (r.x - gv.x) - a * hd.x, hy = (r.y - gv.y) - a * hd.y;  // Heta - tau*Hmdelta (:196)
          q[0] += n.x * gv.x + n.y * gv.y;
          q[1] += n.x * hx + n.y * hy;
          q[2] += n.x * n.x + n.y * n.y;
        } else {
          double2 rn;
          rn.x = r.x - a * hd.x;  // :238
          rn.y = r.y - a * hd.y;
          st2(v.r + i, rn);
          q[0] += n.x * gv.x + n.y * gv.y;
          q[1] += n.x * (rn.x - gv.x) + n.y * (rn.y - gv.y);
          q[2] += rn.x * rn.x + rn.y * rn.y;
          q[3] += n.x * n.x + n.y * n.y;
          const double2 y = ld2(Y + i);
          s1 += y.x * rn.x + y.y * rn.y;
          s2 += y.x * d.x + y.y * d.y;
        }
      }
    }
    if (branch == 0) {
      s1 = group_sum<GS>(s1, mask);
      s2 = group_sum<GS>(s2, mask);
      if (gl == 0) {
        t1[row] = s1;
        t2[row] = s2;
      }
    }
  }
  double tot[6];
  if (grid_sum_last<6>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0)
      for (int k = 0; k < 6; ++k) pack[k] = tot[k];
  }
}
__global__ void k_col_after_update(RtrState* st, const double* pack, cudaGraphConditionalHandle cond, int use_cond) {
  if (st->stop != 0) return;
  tcg_after_update(st, pack, cond, use_cond);
}
// mdelta = (r + beta*mdelta) - Y .* (t1 + beta*t2)   (tCG.m:273,283)
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_col_tcg_dir(VecPtrs v, RtrState* st, const double* __restrict__ pack, int64_t nrows, int ld) {
  if (st->stop != 0) return;
  const double beta = st->beta;
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const double* __restrict__ t1 = pack + 8;
  const double* __restrict__ t2 = pack + 8 + nrows;
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    const double dot = t1[row] + beta * t2[row];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = gl + GS * k;
      if (c < nvec) {
        const size_t i = base + 2 * c;
        const double2 r = ld2(v.r + i), d = ld2(v.d + i), y = ld2(Y + i);
        double2 dn;
        dn.x = (r.x + beta * d.x) - y.x * dot;
        dn.y = (r.y + beta * d.y) - y.y * dot;
        st2(v.d + i, dn);
      }
    }
  }
}

// update pass + ONE all-reduce + scalar logic + new direction of one split tCG iteration
int msdp_col_tcg_update_dir(manisdp_handle* h, cudaGraphConditionalHandle cond, int use_cond) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_tcg_update<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->partials, h->col_pack, h->nloc, ld);
  });
  KERNEL_CHECK(h);
  MSDP_TRY(col_allreduce(h, h->col_pack, 8 + 2 * h->nloc));
  k_col_after_update<<<1, 1, 0, h->stream>>>(h->st, h->col_pack, cond, use_cond);
  KERNEL_CHECK(h);
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    k_col_tcg_dir<GS, VPL><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->col_pack, h->nloc, ld);
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_col_allreduce_tmp(manisdp_handle* h, int count) { return col_allreduce(h, h->st->tmp, count); }

// scalar tail of a cost+grad call: tmp[0] = sum(eG) (all-reduced), tmp[1] = |G|^2  ->  what spmm_tail<EPI_COSTGRAD> does
__global__ void k_col_cg_scalar(RtrState* st, int mode) {
  const double f = 0.5 * st->tmp[0];  // ManiSDP_onlyunitdiag.m:120
  const double g2 = st->tmp[1];
  st->tmp[0] = f;
  if (mode == CG_INIT) {
    st->fx = f;
    st->gradnorm2 = g2;
  } else if (mode == CG_TR) {
    st->fprop = f;
    st->gradnorm2_prop = g2;
    tr_decide(st);
  }
}
int msdp_col_cg_scalar(manisdp_handle* h, int cg_mode) {
  k_col_cg_scalar<<<1, 1, 0, h->stream>>>(h->st, cg_mode);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// ---- split / merge -------------------------------------------------------------------------------------------------------
// dst (n x ldd) = columns [c0, c0 + pl) of src (n x lds, psrc valid columns), zero past psrc and in the ld padding
__global__ void k_col_take(const double* __restrict__ src, int lds, int psrc, double* __restrict__ dst, int ldd, int pl,
                           int c0, int64_t nrows) {
  const int64_t total = nrows * ldd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ldd;
    const int j = (int)(i - row * ldd);
    const int c = c0 + j;
    dst[i] = (j < pl && c < psrc) ? src[(size_t)row * lds + c] : 0.0;
  }
}
// dst (n x ldd, G*pl valid columns) <- G slices src[q] (n x lds, pl valid columns each), slice q at columns [q*pl, ...)
__global__ void k_col_put(const double* __restrict__ src, int lds, int pl, int G, double* __restrict__ dst, int ldd,
                          int64_t nrows) {
  const int64_t total = nrows * ldd;
  const size_t slice = (size_t)nrows * lds;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ldd;
    const int c = (int)(i - row * ldd);
    const int q = c / pl, j = c - q * pl;
    dst[i] = (q < G) ? src[(size_t)q * slice + (size_t)row * lds + j] : 0.0;
  }
}

static int flat_blocks(const manisdp_handle* h, int64_t total) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->num_sms * 8, (total + 255) / 256));
}

int msdp_col_split(manisdp_handle* h) {
  if (!h->col_mode) return msdp_fail(h, MANISDP_E_STATE, "col_split: not a column-sharded handle");
  if (h->col_split) return MANISDP_OK;
  if (h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "col_split: set_Y / rand_Y first");
  const int G = h->cworld;
  const int pfull = (int)h->p, ldf = (int)h->ld;
  const int pl = (pfull + G - 1) / G;
  double* tmp = nullptr;
  const size_t bytes = (size_t)h->nloc * ldf * sizeof(double);
  CUDA_TRY(h, cudaMalloc((void**)&tmp, bytes));
  cudaError_t e = cudaMemcpyAsync(tmp, h->Ybuf[h->pt], bytes, cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  int rc = MANISDP_OK;
  if (e == cudaSuccess) rc = msdp_resize(h, pl);
  if (e == cudaSuccess && rc == MANISDP_OK) {
    k_col_take<<<flat_blocks(h, h->nloc * h->ld), 256, 0, h->stream>>>(tmp, ldf, pfull, h->Ybuf[h->pt], (int)h->ld, pl,
                                                                        h->crank * pl, h->nloc);
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(tmp);
  MSDP_TRY(rc);
  CUDA_TRY(h, e);
  h->col_split = 1;
  h->col_pfull = pfull;
  h->cache_valid = h->grad_valid = 0;
  h->y_version++;
  return MANISDP_OK;
}

int msdp_col_merge(manisdp_handle* h) {
  if (!h->col_mode) return msdp_fail(h, MANISDP_E_STATE, "col_merge: not a column-sharded handle");
  if (!h->col_split) return MANISDP_OK;
  const int G = h->cworld;
  const int pl = (int)h->p, ldl = (int)h->ld;
  const size_t slice = (size_t)h->nloc * ldl;
  double* all = nullptr;
  CUDA_TRY(h, cudaMalloc((void**)&all, slice * G * sizeof(double)));
  int rc = MANISDP_OK;
  cudaError_t e = cudaSuccess;
  if (G > 1) {
    ncclResult_t r = g_cn.AllGather(h->Ybuf[h->pt], all, slice, ncclDouble, (ncclComm_t)h->col_comm, h->stream);
    if (r != ncclSuccess) rc = msdp_fail(h, MANISDP_E_NCCL, std::string("col_merge all-gather: ") + g_cn.GetErrorString(r));
  } else {
    e = cudaMemcpyAsync(all, h->Ybuf[h->pt], slice * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  h->col_split = 0;  // (resize below must not see a split handle)
  if (rc == MANISDP_OK && e == cudaSuccess) rc = msdp_resize(h, (int64_t)G * pl);
  if (rc == MANISDP_OK && e == cudaSuccess) {
    k_col_put<<<flat_blocks(h, h->nloc * h->ld), 256, 0, h->stream>>>(all, ldl, pl, G, h->Ybuf[h->pt], (int)h->ld,
                                                                       h->nloc);
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(all);
  MSDP_TRY(rc);
  CUDA_TRY(h, e);
  h->cache_valid = h->grad_valid = 0;
  h->y_version++;
  return MANISDP_OK;
}
