// api.cu -- extern "C" surface of libmanisdp_b200.so (include/manisdp_b200.h): lifetime, state transfer, closures.
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include "affine.h"
#include "dist.h"
#include "kernels.cuh"
#include "rowops.cuh"

static thread_local std::string g_create_error;

int msdp_fail(manisdp_handle* h, int code, const std::string& msg) {
  if (h)
    h->err = msg;
  else
    g_create_error = msg;
  return code;
}

// ---- small kernels: layout conversion, random init -----------------------------------------------------------------
// dst (n x ld rows, zero padded) <- src (n x p column-major)
__global__ void k_cols_to_rows(const double* src, double* dst, int64_t n, int64_t p, int64_t ld) {
  const int64_t total = n * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / ld, c = i % ld;
    dst[i] = (c < p) ? src[c * n + r] : 0.0;
  }
}
__global__ void k_rows_to_cols(const double* src, double* dst, int64_t n, int64_t p, int64_t ld) {
  const int64_t total = n * p, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t c = i / n, r = i % n;
    dst[i] = src[r * ld + c];
  }
}

// Philox4x32-10 (Salmon et al. 2011), counter = element index, key = seed
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ double philox_normal(uint64_t idx, uint64_t seed) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint64_t a = ((uint64_t)c[0] << 32) | c[1], b = ((uint64_t)c[2] << 32) | c[3];
  const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
  const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);                            // Box-Muller
}
__global__ void k_randn_rows(double* dst, int64_t nloc, int64_t p, int64_t ld, int64_t row0, uint64_t seed) {
  const int64_t total = nloc * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / ld, c = i % ld;
    dst[i] = (c < p) ? philox_normal((uint64_t)((row0 + r) * p + c), seed) : 0.0;
  }
}

__global__ void k_set_pt(RtrState* st, int pt) { st->pt = pt; st->ticket = 0u; }

static int grid_for(const manisdp_handle* h, int64_t total) {
  int64_t nb = (total + MSDP_THREADS - 1) / MSDP_THREADS;
  const int64_t cap = (int64_t)h->num_sms * 8;
  return (int)std::max<int64_t>(1, std::min(nb, cap));
}

// ---- allocation ----------------------------------------------------------------------------------------------------
static int64_t rows_alloc(const manisdp_handle* h) { return msdp_rows_per_rank(h->n, h->world); }

int msdp_scratch(manisdp_handle* h, int slot, size_t bytes, void** out) {
  if (bytes > h->scratch_cap[slot]) {
    if (h->scratch[slot]) {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      cudaFree(h->scratch[slot]);
      h->scratch[slot] = nullptr;
      h->scratch_cap[slot] = 0;
    }
    const size_t cap = bytes + bytes / 2 + 4096;
    CUDA_TRY(h, cudaMalloc(&h->scratch[slot], cap));
    h->scratch_cap[slot] = cap;
  }
  *out = h->scratch[slot];
  return MANISDP_OK;
}

int msdp_resize(manisdp_handle* h, int64_t p) {
  if (p < 1) return msdp_fail(h, MANISDP_E_ARG, "p must be >= 1");
  const int64_t ld = 4 * ((p + 3) / 4);
  const int64_t max_ld = (h->kind == MANISDP_ONLYUNITDIAG || h->kind == MANISDP_MULTIBLOCK) ? MSDP_MAX_LD : MSDP_MAX_LD_AFFINE;
  if (ld > max_ld)
    return msdp_fail(h, MANISDP_E_ARG, max_ld == MSDP_MAX_LD ? "factor width p > 512 is not supported for this kind"
                                                            : "factor width p > 1024 is not supported");
  const size_t need = (size_t)rows_alloc(h) * (size_t)ld;
  if (need > h->cap_elems) {
    // grow with headroom: the outer loop appends up to delta columns per iteration
    const int64_t ld_cap = std::min<int64_t>(max_ld, 4 * ((ld + ld / 2 + 16 + 3) / 4));
    const size_t cap = (size_t)rows_alloc(h) * (size_t)ld_cap;
    double** arrs[] = {&h->Ybuf[0], &h->Ybuf[1], &h->Gbuf[0], &h->Gbuf[1], &h->eta[0], &h->eta[1],
                       &h->r,       &h->d,       &h->Hd,      &h->Uslot,   &h->Hslot};
    MSDP_TRY(msdp_dist_ipc_release(h));  // peers unmap the old arrays before they are freed
    msdp_invalidate_graph(h);
    // From here until every allocation has succeeded the handle holds NO factor: a failure (e.g. out of memory) must
    // leave it in a state where later calls fail cleanly ("no factor set") instead of launching kernels on freed
    // pointers with the old capacity.
    h->cap_elems = 0;
    h->p = 0;
    h->ld = 0;
    h->cache_valid = h->grad_valid = 0;
    for (double** a : arrs) {
      if (*a) cudaFree(*a);
      *a = nullptr;
    }
    if (h->gatherbuf) cudaFree(h->gatherbuf);
    h->gatherbuf = nullptr;
    cudaError_t aerr = cudaSuccess;
    for (double** a : arrs) {
      aerr = cudaMalloc((void**)a, cap * sizeof(double));
      if (aerr == cudaSuccess) aerr = cudaMemsetAsync(*a, 0, cap * sizeof(double), h->stream);
      if (aerr != cudaSuccess) break;
    }
    if (aerr == cudaSuccess && h->world > 1) {
      aerr = cudaMalloc((void**)&h->gatherbuf, cap * h->world * sizeof(double));
      if (aerr == cudaSuccess) aerr = cudaMemsetAsync(h->gatherbuf, 0, cap * h->world * sizeof(double), h->stream);
    }
    if (aerr != cudaSuccess) {
      for (double** a : arrs) {
        if (*a) cudaFree(*a);
        *a = nullptr;
      }
      if (h->gatherbuf) cudaFree(h->gatherbuf);
      h->gatherbuf = nullptr;
      cudaGetLastError();
      return msdp_fail(h, MANISDP_E_CUDA, std::string("resize: allocation of the work arrays failed: ") +
                                              cudaGetErrorString(aerr));
    }
    h->cap_elems = cap;
    msdp_invalidate_graph(h);
    MSDP_TRY(msdp_dist_ipc_refresh(h));
  }
  if (h->bm_B > 0 && ld > 32 && ld <= 64) {  // partial rows of the block-major product (spmm.cu: launch_bm)
    const size_t want = (size_t)h->bm_B * (size_t)rows_alloc(h) * 64;
    if (h->bm_part_cap < want) {
      if (h->bm_part) cudaFree(h->bm_part);
      h->bm_part = nullptr;
      CUDA_TRY(h, cudaMalloc((void**)&h->bm_part, want * sizeof(double)));
      h->bm_part_cap = want;
      h->bm_part_ld = -1;
      msdp_invalidate_graph(h);
    }
    // k_bm_finish sums ALL blocks of a row: (block, row) pairs without entries are never written by a pass and must
    // read as exact zeros, so the buffers are cleared whenever the row length (= the layout) changes
    if (h->bm_part_ld != ld) {
      CUDA_TRY(h, cudaMemsetAsync(h->bm_part, 0, h->bm_part_cap * sizeof(double), h->stream));
      h->bm_part_ld = ld;
    }
  }
  if (p != h->p) msdp_invalidate_graph(h);
  h->y_version++;
  h->p = p;
  h->ld = ld;
  h->cache_valid = 0;
  h->grad_valid = 0;
  return MANISDP_OK;
}

static double* slot_ptr(manisdp_handle* h, int slot) {
  switch (slot) {
    case MANISDP_SLOT_Y: return h->Ybuf[h->pt];
    case MANISDP_SLOT_YPROP: return h->Ybuf[h->pt ^ 1];
    case MANISDP_SLOT_G: return h->Gbuf[h->pt];
    case MANISDP_SLOT_ETA: return h->eta[h->st_host ? h->st_host->eta_cur : 0];
    case MANISDP_SLOT_R: return h->r;
    case MANISDP_SLOT_D: return h->d;
    case MANISDP_SLOT_HD: return h->Hd;
    case MANISDP_SLOT_U: return h->Uslot;
    case MANISDP_SLOT_H: return h->Hslot;
  }
  return nullptr;
}

// host (nloc x p, layout) -> device rows array
static int upload_rows(manisdp_handle* h, double* dst, const double* src, int32_t layout) {
  const int64_t n = h->nloc, p = h->p, ld = h->ld;
  CUDA_TRY(h, cudaMemsetAsync(dst, 0, (size_t)rows_alloc(h) * ld * sizeof(double), h->stream));
  if (layout == MANISDP_LAYOUT_ROWS) {
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, ld * sizeof(double), src, p * sizeof(double), p * sizeof(double), n,
                                  cudaMemcpyHostToDevice, h->stream));
  } else {
    double* stage = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&stage, (size_t)n * p * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(stage, src, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      k_cols_to_rows<<<grid_for(h, n * ld), MSDP_THREADS, 0, h->stream>>>(stage, dst, n, p, ld);
      h->launches++;
      e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(stage);
    CUDA_TRY(h, e);
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MANISDP_OK;
}

static int download_rows(manisdp_handle* h, const double* src, double* dst, int32_t layout) {
  const int64_t n = h->nloc, p = h->p, ld = h->ld;
  if (layout == MANISDP_LAYOUT_ROWS) {
    CUDA_TRY(h, cudaMemcpy2DAsync(dst, p * sizeof(double), src, ld * sizeof(double), p * sizeof(double), n,
                                  cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  } else {
    double* stage = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&stage, (size_t)n * p * sizeof(double)));
    k_rows_to_cols<<<grid_for(h, n * p), MSDP_THREADS, 0, h->stream>>>(src, stage, n, p, ld);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(dst, stage, (size_t)n * p * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(stage);
    CUDA_TRY(h, e);
  }
  return MANISDP_OK;
}

// ---- create / destroy ----------------------------------------------------------------------------------------------
// Block-major copy of the row lists (rp, ci, val) for spmm.cu:launch_bm.  Operand rows are cut into B column blocks
// (about 2 x the L2 target at the nominal width p = 64, at most 8 blocks); the entries are stored block by block, inside
// a block row by row in their original order, as (col, val, row); every block is cut into row-aligned chunks of about
// BM_CHUNK entries (one chunk = the work of one warp).  All index work is integer and exact.
#define BM_CHUNK 224
static int build_block_major(manisdp_handle* h, const std::vector<int>& rp, const std::vector<int>& ci,
                             const double* val, int64_t nrows) {
  const int64_t nnz = (int64_t)ci.size();
  const int64_t ncols_op = h->n;  // operand rows = global columns of C
  int B;
  if (h->bm_mode == 2 && ncols_op * 512 < 4 * h->spmm_l2_target)
    B = 3;  // forced on a small instance (tests): still several blocks
  else
    B = (int)std::min<int64_t>(8, std::max<int64_t>(2, (ncols_op * 512 + 2 * h->spmm_l2_target - 1) /
                                                           (2 * h->spmm_l2_target)));
  const int64_t jrows = (ncols_op + B - 1) / B;
  std::vector<int64_t> start((size_t)B + 1, 0);
  for (int64_t e = 0; e < nnz; ++e) start[(size_t)(ci[(size_t)e] / jrows) + 1]++;
  for (int b = 0; b < B; ++b) start[(size_t)b + 1] += start[(size_t)b];
  std::vector<int> ecol((size_t)nnz), erow((size_t)nnz);
  std::vector<double> ev((size_t)nnz);
  std::vector<int64_t> pos(start.begin(), start.end() - 1);
  for (int64_t i = 0; i < nrows; ++i)
    for (int e = rp[(size_t)i]; e < rp[(size_t)i + 1]; ++e) {
      const int b = (int)(ci[(size_t)e] / jrows);
      const int64_t q = pos[(size_t)b]++;
      ecol[(size_t)q] = ci[(size_t)e];
      erow[(size_t)q] = (int)i;
      ev[(size_t)q] = val[e];
    }
  std::vector<int> cptr;
  h->bm_chunk_off.assign((size_t)B + 1, 0);
  for (int b = 0; b < B; ++b) {
    h->bm_chunk_off[(size_t)b] = (int)cptr.size();
    int64_t e = start[(size_t)b];
    const int64_t end = start[(size_t)b + 1];
    while (e < end) {
      cptr.push_back((int)e);
      int64_t f = std::min(end, e + BM_CHUNK);
      while (f < end && erow[(size_t)f] == erow[(size_t)f - 1]) ++f;  // a row never straddles two chunks
      e = f;
    }
    cptr.push_back((int)end);  // closes the last chunk of the block (a block's chunk table has nchunks + 1 entries)
  }
  h->bm_chunk_off[(size_t)B] = (int)cptr.size();
  CUDA_TRY(h, cudaMalloc((void**)&h->bm_col, (size_t)nnz * sizeof(int)));
  CUDA_TRY(h, cudaMalloc((void**)&h->bm_row, (size_t)nnz * sizeof(int)));
  CUDA_TRY(h, cudaMalloc((void**)&h->bm_val, (size_t)nnz * sizeof(double)));
  CUDA_TRY(h, cudaMalloc((void**)&h->bm_chunk, cptr.size() * sizeof(int)));
  CUDA_TRY(h, cudaMemcpy(h->bm_col, ecol.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->bm_row, erow.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->bm_val, ev.data(), (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->bm_chunk, cptr.data(), cptr.size() * sizeof(int), cudaMemcpyHostToDevice));
  h->bm_B = B;
  return MANISDP_OK;
}

static int upload_csr_from_csc(manisdp_handle* h, Csr& out, const uint64_t* jc, const uint64_t* ir, const double* pr,
                               int64_t ncols, int64_t nrows_global) {
  // The column lists of C are exactly the row lists the kernels need: out(j,:) = sum_i C(i,j) Y(i,:) is the
  // reference's (Y*C)(:,j) (ManiSDP_onlyunitdiag.m:118) whether or not C is symmetric.
  const uint64_t nnz = jc[ncols] - jc[0];
  if (nnz >= (1ull << 31) || nrows_global >= (1ll << 31))
    return msdp_fail(h, MANISDP_E_ARG, "C: nnz and n must be < 2^31");
  std::vector<int> rp((size_t)ncols + 1), ci((size_t)nnz);
  const uint64_t base = jc[0];
  for (int64_t j = 0; j <= ncols; ++j) rp[(size_t)j] = (int)(jc[j] - base);
  for (uint64_t e = 0; e < nnz; ++e) {
    if (ir[base + e] >= (uint64_t)nrows_global) return msdp_fail(h, MANISDP_E_ARG, "C: row index out of range");
    ci[(size_t)e] = (int)ir[base + e];
  }
  // locality statistics for the column-blocking decision of spmm.cu
  {
    const char* em = getenv("MANISDP_SPMM_BLOCK");
    if (em) h->spmm_block_mode = atoi(em);
    const char* eb = getenv("MANISDP_SPMM_BULK");
    if (eb) h->spmm_use_bulk = atoi(eb);
    const char* en = getenv("MANISDP_SPMM_NARROW");
    if (en) h->spmm_narrow = atoi(en);
    const char* el = getenv("MANISDP_L2_TARGET_MB");
    if (el && atoi(el) > 0) h->spmm_l2_target = (int64_t)atoi(el) << 20;
    const int64_t window = h->spmm_l2_target / (64 * 8);  // operand rows per L2 window at the nominal width p = 64
    bool sorted = true;
    uint64_t far = 0;
    for (int64_t j = 0; j < ncols; ++j) {
      const int64_t grow = h->row_begin + j;
      for (int e = rp[(size_t)j]; e < rp[(size_t)j + 1]; ++e) {
        if (e > rp[(size_t)j] && ci[(size_t)e] <= ci[(size_t)e - 1]) sorted = false;
        const int64_t dlt = (int64_t)ci[(size_t)e] - grow;
        if (dlt > window || dlt < -window) ++far;
      }
    }
    h->C_sorted = sorted ? 1 : 0;
    {  // eligibility of the batched low-degree kernel (spmm.cu: k_spmm_lowdeg, LB_CAP = 320 staged entries per batch)
      int64_t worst = 0, maxdeg = 0;
      for (int64_t j = 0; j < ncols; ++j) maxdeg = std::max<int64_t>(maxdeg, (int64_t)rp[(size_t)j + 1] - rp[(size_t)j]);
      h->C_maxdeg = (int)std::min<int64_t>(maxdeg, 1 << 30);
      for (int64_t j = 0; j < ncols; j += 32) {
        const int64_t j1 = std::min<int64_t>(ncols, j + 32);
        worst = std::max<int64_t>(worst, (int64_t)rp[(size_t)j1] - rp[(size_t)j]);
      }
      h->C_lowdeg = (ncols > 0 && worst <= 320 && (double)nnz <= 8.0 * (double)ncols) ? 1 : 0;
      const char* eld = getenv("MANISDP_SPMM_LOWDEG");
      if (eld) h->spmm_lowdeg = atoi(eld);
    }
    h->C_far_fraction = nnz ? (double)far / (double)nnz : 0.0;
    uint64_t remote = 0;
    for (uint64_t e = 0; e < nnz; ++e)
      remote += (ci[(size_t)e] < h->row_begin || ci[(size_t)e] >= h->row_end) ? 1 : 0;
    h->C_remote_fraction = nnz ? (double)remote / (double)nnz : 0.0;
    const char* eg = getenv("MANISDP_PEER_GATHER_MAX");
    if (eg) h->peer_gather_max_remote = atof(eg);
  }
  {
    const char* ebm = getenv("MANISDP_SPMM_BM");
    if (ebm) h->bm_mode = atoi(ebm);
    // auto: single GPU, no locality, operand at the nominal width p = 64 at least twice the L2 target
    const bool want = h->world == 1 && nnz > 0 &&
                      (h->bm_mode == 2 || (h->bm_mode == 1 && h->C_far_fraction > 0.5 &&
                                           ncols * 512 >= 4 * h->spmm_l2_target));
    if (want) MSDP_TRY(build_block_major(h, rp, ci, pr + base, ncols));
  }
  out.nrows = ncols;
  out.nnz = (int64_t)nnz;
  CUDA_TRY(h, cudaMalloc((void**)&out.rowptr, rp.size() * sizeof(int)));
  CUDA_TRY(h, cudaMalloc((void**)&out.col, std::max<size_t>(1, ci.size()) * sizeof(int)));
  CUDA_TRY(h, cudaMalloc((void**)&out.val, std::max<size_t>(1, (size_t)nnz) * sizeof(double)));
  CUDA_TRY(h, cudaMemcpy(out.rowptr, rp.data(), rp.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(out.col, ci.data(), ci.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(out.val, pr + base, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
  return MANISDP_OK;
}

static int create_impl(manisdp_handle* h, const manisdp_problem* pb) {
  NvtxRange nvtx_range("manisdp:create");
  if (pb->kind < MANISDP_ONLYUNITDIAG || pb->kind > MANISDP_DUAL_UNITDIAG) return msdp_fail(h, MANISDP_E_ARG, "bad kind");
  if (pb->n < 1) return msdp_fail(h, MANISDP_E_ARG, "n must be >= 1");
  h->kind = pb->kind;
  h->mf = (pb->kind == MANISDP_UNITTRACE) ? MF_SPHERE : (pb->kind == MANISDP_GENERAL ? MF_EUCLID : MF_OBLIQUE);
  h->device = pb->device;
  h->n = pb->n;
  h->m = pb->m;
  const bool col_layout = (pb->shard_layout == MANISDP_SHARD_COLS);
  h->world = (pb->world > 1 && !col_layout) ? pb->world : 1;
  h->rank = h->world > 1 ? pb->rank : 0;
  int ndev = 0;
  CUDA_TRY(h, cudaGetDeviceCount(&ndev));
  if (ndev < 1) return msdp_fail(h, MANISDP_E_CUDA, "no CUDA device: this engine has no CPU fallback");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, h->device));
  if (prop.major < 10) return msdp_fail(h, MANISDP_E_CUDA, "an sm_100 (B200) device is required");
  h->num_sms = prop.multiProcessorCount;
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaEventCreate(&h->ev0));
  CUDA_TRY(h, cudaEventCreate(&h->ev1));
  if (h->world > 1) {
    if (h->kind != MANISDP_ONLYUNITDIAG)
      return msdp_fail(h, MANISDP_E_ARG, "row sharding is implemented for ONLYUNITDIAG (SURVEY 8e: configs 2-4 stay on one GPU)");
    const int64_t rpr = msdp_rows_per_rank(h->n, h->world);
    h->row_begin = std::min<int64_t>(h->n, rpr * h->rank);
    h->row_end = std::min<int64_t>(h->n, h->row_begin + rpr);
    if (pb->row_begin != h->row_begin || pb->row_end != h->row_end)
      return msdp_fail(h, MANISDP_E_ARG, "sharded handle: rows must be [rank*ceil(n/world), ...)");
    MSDP_TRY(msdp_dist_init(h, pb->nccl_unique_id));
  } else {
    h->row_begin = 0;
    h->row_end = h->n;
  }
  h->nloc = h->row_end - h->row_begin;
  if (col_layout) {
    if (h->kind != MANISDP_ONLYUNITDIAG)
      return msdp_fail(h, MANISDP_E_ARG, "column sharding is implemented for ONLYUNITDIAG (SURVEY 8e: configs 2-4 stay on one GPU)");
    if (pb->row_begin != 0 || (pb->row_end != 0 && pb->row_end != pb->n))
      return msdp_fail(h, MANISDP_E_ARG, "column-sharded handle: every rank holds all rows (row_begin = 0, row_end = n)");
    MSDP_TRY(msdp_col_init(h, pb->nccl_unique_id, pb->world, pb->rank));
  }
  CUDA_TRY(h, cudaMalloc((void**)&h->st, sizeof(RtrState)));
  CUDA_TRY(h, cudaMemset(h->st, 0, sizeof(RtrState)));
  {  // per-row manifold switch of the oblique kernels: every row is a unit vector unless a multi-block setup says otherwise
    const long long all_rows = 0x7fffffffffffffffll;
    CUDA_TRY(h, cudaMemcpy(&h->st->nob_rows, &all_rows, sizeof(long long), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(h, cudaMallocHost((void**)&h->st_host, sizeof(RtrState)));
  memset(h->st_host, 0, sizeof(RtrState));
  CUDA_TRY(h, cudaMalloc((void**)&h->partials, sizeof(double) * MSDP_NQ * MSDP_MAX_BLOCKS));
  const size_t nalloc = (size_t)rows_alloc(h);
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(h, cudaMalloc((void**)&h->eG[i], nalloc * sizeof(double)));
    CUDA_TRY(h, cudaMemset(h->eG[i], 0, nalloc * sizeof(double)));
  }
  CUDA_TRY(h, cudaMalloc((void**)&h->zdiag, nalloc * sizeof(double)));
  CUDA_TRY(h, cudaMemset(h->zdiag, 0, nalloc * sizeof(double)));
  if (h->kind == MANISDP_ONLYUNITDIAG) {
    if (!pb->C_jc || !pb->C_ir || !pb->C_pr) return msdp_fail(h, MANISDP_E_ARG, "ONLYUNITDIAG needs C (CSC)");
    MSDP_TRY(upload_csr_from_csc(h, h->C, pb->C_jc, pb->C_ir, pb->C_pr, h->nloc, h->n));
    h->s_mode = MODE_SPARSE;
    h->a_mode = MODE_NONE;
  } else {
    if (h->kind == MANISDP_MULTIBLOCK) {
      if (pb->world > 1) return msdp_fail(h, MANISDP_E_ARG, "multi-block handles are single-GPU (SURVEY 8e: small blocks, latency-bound)");
      MSDP_TRY(msdp_mb_setup(h, pb));
    }
    if (h->kind == MANISDP_DUAL_UNITDIAG) {
      // the engine sees A~ = D^-1/2 A, b_eff = A~ c and (rebuilt every ADMM step) C_eff = bA + x - sigma*c  (dual.cu)
      if (!pb->At_jc || !pb->At_ir || !pb->At_pr || !pb->b || !pb->c_pr)
        return msdp_fail(h, MANISDP_E_ARG, "dual handles need At, b and c");
      const int64_t m = pb->m;
      const uint64_t nz = pb->At_jc[m], nn = (uint64_t)pb->n * (uint64_t)pb->n;
      std::vector<double> dAAt((size_t)m, 0.0), pr(pb->At_pr, pb->At_pr + nz), beff((size_t)m, 0.0), cd((size_t)nn, 0.0);
      if (pb->c_ir) {
        for (int64_t q = 0; q < pb->c_nnz; ++q) {
          if (pb->c_ir[q] >= nn) return msdp_fail(h, MANISDP_E_ARG, "c: index out of range");
          cd[(size_t)pb->c_ir[q]] += pb->c_pr[q];
        }
      } else {
        if ((uint64_t)pb->c_nnz != nn) return msdp_fail(h, MANISDP_E_ARG, "dense c must have n*n entries");
        std::copy(pb->c_pr, pb->c_pr + nn, cd.begin());
      }
      for (int64_t k = 0; k < m; ++k) {
        double dk = 0.0;
        for (uint64_t e = pb->At_jc[k]; e < pb->At_jc[k + 1]; ++e) dk += pb->At_pr[e] * pb->At_pr[e];
        dAAt[(size_t)k] = pb->dAAt ? pb->dAAt[k] : dk;  // ManiDSDP_unitdiag.m:40
        if (!(dAAt[(size_t)k] > 0.0)) return msdp_fail(h, MANISDP_E_ARG, "dual: diag(A*A') must be positive");
        const double is = 1.0 / sqrt(dAAt[(size_t)k]);
        double acc = 0.0;
        for (uint64_t e = pb->At_jc[k]; e < pb->At_jc[k + 1]; ++e) {
          if (pb->At_ir[e] >= nn) return msdp_fail(h, MANISDP_E_ARG, "At: row index out of range");
          pr[(size_t)e] *= is;
          acc += pr[(size_t)e] * cd[(size_t)pb->At_ir[e]];
        }
        beff[(size_t)k] = acc;
      }
      manisdp_problem q = *pb;
      q.At_pr = pr.data();
      q.b = beff.data();
      q.force_mode = (pb->force_mode & ~2) | 1;  // dense S: C_eff is a dense n x n matrix
      MSDP_TRY(msdp_affine_setup(h, &q));
      MSDP_TRY(msdp_dual_setup(h, pb, dAAt));
      return MANISDP_OK;
    }
    MSDP_TRY(msdp_affine_setup(h, pb));
  }
  return MANISDP_OK;
}

static void free_all(manisdp_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  msdp_invalidate_graph(h);
  msdp_eig_release(h);
  msdp_dist_destroy(h);
  msdp_col_destroy(h);
  msdp_affine_free(h);
  msdp_mb_free(h);
  msdp_dual_free(h);
  double* arrs[] = {h->Ybuf[0], h->Ybuf[1], h->Gbuf[0], h->Gbuf[1], h->eta[0], h->eta[1], h->r, h->d, h->Hd,
                    h->Uslot, h->Hslot, h->gatherbuf, h->eG[0], h->eG[1], h->zdiag, h->partials, h->C.val,
                    h->eigvecs};
  for (double* a : arrs)
    if (a) cudaFree(a);
  if (h->C.rowptr) cudaFree(h->C.rowptr);
  if (h->spmm_bptr) cudaFree(h->spmm_bptr);
  if (h->gemm_ws) cudaFree(h->gemm_ws);
  if (h->owner_bptr) cudaFree(h->owner_bptr);
  for (void* q : h->scratch)
    if (q) cudaFree(q);
  void* bm[] = {h->bm_col, h->bm_row, h->bm_chunk, h->bm_val, h->bm_part};
  for (void* q : bm)
    if (q) cudaFree(q);
  if (h->C.col) cudaFree(h->C.col);
  if (h->st) cudaFree(h->st);
  if (h->st_host) cudaFreeHost(h->st_host);
  if (h->eigvals_host) free(h->eigvals_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" {

int manisdp_version(void) { return 100; }

int manisdp_create(manisdp_t** out, const manisdp_problem* prob) {
  if (!out || !prob) return msdp_fail(nullptr, MANISDP_E_ARG, "null argument");
  *out = nullptr;
  manisdp_handle* h = new (std::nothrow) manisdp_handle();
  if (!h) return msdp_fail(nullptr, MANISDP_E_ARG, "out of host memory");
  int rc = create_impl(h, prob);
  if (rc != MANISDP_OK) {
    g_create_error = h->err;
    free_all(h);
    return rc;
  }
  *out = h;
  return MANISDP_OK;
}

int manisdp_destroy(manisdp_t* h) {
  free_all(h);
  return MANISDP_OK;
}

const char* manisdp_last_error(const manisdp_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int manisdp_set_Y(manisdp_t* h, const double* Y, int64_t p, int32_t layout) {
  if (!h || !Y) return msdp_fail(h, MANISDP_E_ARG, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  // same width: the captured trust-region graphs stay valid (msdp_resize only invalidates them when p changes)
  MSDP_TRY(msdp_resize(h, p));
  return upload_rows(h, h->Ybuf[h->pt], Y, layout);
}

int manisdp_get_Y(manisdp_t* h, double* Y, int32_t layout) {
  if (!h || !Y) return msdp_fail(h, MANISDP_E_ARG, "null argument");
  if (h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return download_rows(h, h->Ybuf[h->pt], Y, layout);
}

int manisdp_get_p(manisdp_t* h, int64_t* p) {
  if (!h || !p) return MANISDP_E_ARG;
  *p = h->p;
  return MANISDP_OK;
}

int manisdp_rand_Y(manisdp_t* h, int64_t p, uint64_t seed) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "rand_Y: the handle is column-split (manisdp_col_merge first)");
  if (!h) return MANISDP_E_ARG;
  CUDA_TRY(h, cudaSetDevice(h->device));
  MSDP_TRY(msdp_resize(h, p));
  double* raw = h->Ybuf[h->pt ^ 1];
  k_randn_rows<<<grid_for(h, h->nloc * h->ld), MSDP_THREADS, 0, h->stream>>>(raw, h->nloc, h->p, h->ld,
                                                                           h->row_begin, seed);
  KERNEL_CHECK(h);
  if (h->mf == MF_EUCLID) {
    CUDA_TRY(h, cudaMemcpyAsync(h->Ybuf[h->pt], raw, (size_t)h->nloc * h->ld * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->stream));
  } else {
    // x ./ norms (ManiSDP_unitdiag.m:194-197 / spherefactory.m:249-254) == retr(x, 0)
    CUDA_TRY(h, cudaMemsetAsync(h->eta[0], 0, (size_t)h->nloc * h->ld * sizeof(double), h->stream));
    if (h->mf == MF_SPHERE && h->world > 1) return msdp_fail(h, MANISDP_E_ARG, "sphere + sharding unsupported");
    MSDP_TRY(msdp_launch_retract(h, raw, h->eta[0], h->Ybuf[h->pt], 0));
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MANISDP_OK;
}

int manisdp_slot_set(manisdp_t* h, int32_t slot, const double* src, int32_t layout) {
  if (!h || !src || h->p <= 0) return msdp_fail(h, MANISDP_E_ARG, "slot_set: bad argument / no factor width");
  double* dst = slot_ptr(h, slot);
  if (!dst) return msdp_fail(h, MANISDP_E_ARG, "bad slot");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (slot == MANISDP_SLOT_Y) {
    h->cache_valid = h->grad_valid = 0;
    h->y_version++;
  }
  return upload_rows(h, dst, src, layout);
}

int manisdp_slot_get(manisdp_t* h, int32_t slot, double* dst, int32_t layout) {
  if (!h || !dst || h->p <= 0) return msdp_fail(h, MANISDP_E_ARG, "slot_get: bad argument / no factor width");
  const double* src = slot_ptr(h, slot);
  if (!src) return msdp_fail(h, MANISDP_E_ARG, "bad slot");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return download_rows(h, src, dst, layout);
}

int manisdp_set_dual(manisdp_t* h, const double* y, double sigma) {
  if (!h) return MANISDP_E_ARG;
  if (h->kind == MANISDP_ONLYUNITDIAG) return msdp_fail(h, MANISDP_E_ARG, "ONLYUNITDIAG has no dual vector");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (y && !h->dual.on) CUDA_TRY(h, cudaMemcpy(h->y, y, (size_t)h->m * sizeof(double), cudaMemcpyHostToDevice));
  if (sigma > 0) h->sigma = sigma;
  if (h->dual.on) h->dual.dirty = 1;  // C_eff = bA + x - sigma*c and k0 depend on sigma (dual.cu)
  h->cache_valid = h->grad_valid = 0;
  return MANISDP_OK;
}

int manisdp_set_sigma(manisdp_t* h, double sigma) { return manisdp_set_dual(h, nullptr, sigma); }

int manisdp_get_dual(manisdp_t* h, double* y, double* sigma) {
  if (!h) return MANISDP_E_ARG;
  if (h->kind == MANISDP_ONLYUNITDIAG) return msdp_fail(h, MANISDP_E_ARG, "ONLYUNITDIAG has no dual vector");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (y) CUDA_TRY(h, cudaMemcpy(y, h->y, (size_t)h->m * sizeof(double), cudaMemcpyDeviceToHost));
  if (sigma) *sigma = h->sigma;
  return MANISDP_OK;
}

// ---- closures ------------------------------------------------------------------------------------------------------
int msdp_ensure_costgrad(manisdp_handle* h) {
  if (h->cache_valid && h->grad_valid) return MANISDP_OK;
  if (h->world > 1) {
    MSDP_TRY(msdp_costgrad_exchange(h, h->pt, h->pt, CG_TR_DEFER));
    MSDP_TRY(msdp_dist_allreduce_tmp(h, 2));
    MSDP_TRY(msdp_dist_finish_init(h));
  } else {
    MSDP_TRY(msdp_costgrad(h, h->pt, CG_INIT));
  }
  MSDP_TRY(msdp_sync_state(h));
  h->cache_valid = h->grad_valid = 1;
  return MANISDP_OK;
}

int manisdp_cost(manisdp_t* h, double* f) {
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "cost: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  h->cache_valid = 0;
  MSDP_TRY(msdp_ensure_costgrad(h));
  if (f) *f = h->st_host->fx;
  return MANISDP_OK;
}

int manisdp_grad(manisdp_t* h, double* gradnorm) {
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "grad: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  MSDP_TRY(msdp_ensure_costgrad(h));
  if (gradnorm) *gradnorm = sqrt(h->st_host->gradnorm2);
  return MANISDP_OK;
}

int manisdp_hess(manisdp_t* h) {
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "hess: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  MSDP_TRY(msdp_ensure_costgrad(h));
  if (msdp_pipeline_ok(h)) {
    MSDP_TRY(msdp_maxcut_hess_pipelined(h, h->Uslot, h->Hslot, 0, TAIL_NONE));
    // peers may read SLOT_U in place (direct peer gathers): nobody returns -- and lets its caller overwrite the slot --
    // before every rank's product has run
    MSDP_TRY(msdp_dist_barrier(h));
  } else {
    if (h->world > 1) MSDP_TRY(msdp_dist_allgather_rows(h, h->Uslot, h->gatherbuf));
    MSDP_TRY(msdp_hess_dir(h, h->Uslot, h->Hslot, TAIL_NONE));
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->hv_total += 1;
  return MANISDP_OK;
}

int manisdp_hess_bench(manisdp_t* h, int32_t reps, double* ms_per_hv) {
  if (!h || h->p <= 0 || reps < 1) return msdp_fail(h, MANISDP_E_STATE, "hess_bench: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  MSDP_TRY(msdp_ensure_costgrad(h));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < reps; ++i) {
    if (msdp_pipeline_ok(h)) {
      MSDP_TRY(msdp_maxcut_hess_pipelined(h, h->Uslot, h->Hslot, 0, TAIL_NONE));
    } else {
      if (h->world > 1) MSDP_TRY(msdp_dist_allgather_rows(h, h->Uslot, h->gatherbuf));
      MSDP_TRY(msdp_hess_dir(h, h->Uslot, h->Hslot, TAIL_NONE));
    }
  }
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (ms_per_hv) *ms_per_hv = (double)ms / reps;
  h->hv_total += reps;
  return MANISDP_OK;
}

int manisdp_retract(manisdp_t* h, int32_t eta_slot, int32_t dst_slot) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "retract: the handle is column-split (manisdp_col_merge first)");
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "retract: no factor set");
  double *e = slot_ptr(h, eta_slot), *dst = slot_ptr(h, dst_slot);
  if (!e || !dst || dst == h->Ybuf[h->pt]) return msdp_fail(h, MANISDP_E_ARG, "retract: bad slots");
  CUDA_TRY(h, cudaSetDevice(h->device));
  MSDP_TRY(msdp_launch_retract(h, h->Ybuf[h->pt], e, dst, 0));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MANISDP_OK;
}

int manisdp_project(manisdp_t* h, int32_t src_slot, int32_t dst_slot) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "project: the handle is column-split (manisdp_col_merge first)");
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "project: no factor set");
  double *s = slot_ptr(h, src_slot), *dst = slot_ptr(h, dst_slot);
  if (!s || !dst || dst == h->Ybuf[h->pt]) return msdp_fail(h, MANISDP_E_ARG, "project: bad slots");
  CUDA_TRY(h, cudaSetDevice(h->device));
  MSDP_TRY(msdp_launch_project(h, h->Ybuf[h->pt], s, dst));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MANISDP_OK;
}

int manisdp_tr_solve(manisdp_t* h, const manisdp_tr_options* opts, manisdp_tr_info* info) {
  if (!h) return MANISDP_E_ARG;
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_tr_solve(h, opts, info);
}

int manisdp_tr_log(manisdp_t* h, manisdp_tr_iter* buf, int32_t cap, int32_t* count) {
  if (!h || !count) return MANISDP_E_ARG;
  const int32_t nrec = (int32_t)std::min<size_t>(h->log.size(), (size_t)std::max(cap, 0));
  if (buf)
    for (int32_t i = 0; i < nrec; ++i) buf[i] = h->log[(size_t)i];
  *count = buf ? nrec : (int32_t)h->log.size();
  return MANISDP_OK;
}

int manisdp_kkt(manisdp_t* h, int32_t delta, double eig_tol, int32_t update_dual, manisdp_kkt_info* out) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "kkt: the handle is column-split (manisdp_col_merge first)");
  if (!h || !out) return MANISDP_E_ARG;
  if (h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "kkt: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_kkt(h, delta, eig_tol, update_dual, out);
}

int manisdp_get_eigs(manisdp_t* h, double* vals, double* vecs, int32_t cap) {
  if (!h) return MANISDP_E_ARG;
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int k = std::min<int>(cap, h->eig_k);
  if (vals)
    for (int i = 0; i < k; ++i) vals[i] = h->eigvals_host[i];
  if (vecs && k > 0) {
    CUDA_TRY(h, cudaMemcpy2D(vecs, (size_t)k * sizeof(double), h->eigvecs, (size_t)h->eig_kld * sizeof(double),
                             (size_t)k * sizeof(double), (size_t)h->nloc, cudaMemcpyDeviceToHost));
  }
  return MANISDP_OK;
}

int manisdp_rank_cut(manisdp_t* h, double theta, int32_t apply, int64_t* r, int64_t* p_new) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "rank_cut: the handle is column-split (manisdp_col_merge first)");
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "rank_cut: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_rank_cut(h, theta, apply, r, p_new);
}

int manisdp_escape(manisdp_t* h, int32_t nne, double alpha, int32_t line_search) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "escape: the handle is column-split (manisdp_col_merge first)");
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "escape: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_escape(h, nne, alpha, line_search);
}

int manisdp_get_stats(manisdp_t* h, manisdp_stats* out) {
  if (!h || !out) return MANISDP_E_ARG;
  memset(out, 0, sizeof(*out));
  out->n = h->n;
  out->n_local = h->nloc;
  out->m = h->m;
  out->p = h->p;
  out->ld = h->ld;
  out->nnzC = h->C.nnz;
  out->nnzA = h->a_mode == MODE_DENSE ? h->Ad.nnz : h->As.nnz;
  out->kind = h->kind;
  out->s_mode = h->s_mode;
  out->a_mode = h->a_mode;
  out->rank = h->rank;
  out->world = h->world;
  out->hv_total = h->hv_total;
  out->launches_total = h->launches;
  const double n = (double)h->nloc, p = (double)h->p, nn = (double)h->n * (double)h->n;
  if (h->kind == MANISDP_ONLYUNITDIAG) {
    out->bytes_per_hv = 12.0 * h->C.nnz + 4.0 * (n + 1) + 24.0 * n * p + 8.0 * n;
    out->flops_per_hv = 2.0 * h->C.nnz * p + 5.0 * n * p;
  } else if (h->a_mode == MODE_DENSE) {
    out->bytes_per_hv = 40.0 * nn + 24.0 * out->nnzA + 16.0 * h->m + 24.0 * n * p + 8.0 * n;
    out->flops_per_hv = 6.0 * nn * p + 4.0 * out->nnzA + 6.0 * n * p;
  } else {
    const double nnzS = (h->s_mode == MODE_DENSE) ? nn : (double)h->C.nnz;
    out->bytes_per_hv = 32.0 * out->nnzA + 16.0 * h->m + (h->s_mode == MODE_DENSE ? 8.0 : 12.0) * nnzS + 40.0 * n * p;
    out->flops_per_hv = 2.0 * p * (2.0 * out->nnzA + nnzS);
  }
  return MANISDP_OK;
}

}  // extern "C"

extern "C" int manisdp_line_search(manisdp_t* h, double* alpha) {
  if (h && h->col_split) return msdp_fail(h, MANISDP_E_STATE, "line_search: the handle is column-split (manisdp_col_merge first)");
  if (!h || h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "line_search: no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_line_search(h, alpha);
}

extern "C" int manisdp_col_split(manisdp_t* h) {
  if (!h) return MANISDP_E_ARG;
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_col_split(h);
}
extern "C" int manisdp_col_merge(manisdp_t* h) {
  if (!h) return MANISDP_E_ARG;
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_col_merge(h);
}

extern "C" int manisdp_get_index_split(manisdp_t* h, int64_t* i, int64_t* j, int64_t cap, int64_t* count) {
  if (!h) return MANISDP_E_ARG;
  if (h->kind == MANISDP_ONLYUNITDIAG) return msdp_fail(h, MANISDP_E_ARG, "ONLYUNITDIAG has no constraint matrix At");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return msdp_affine_index_split(h, i, j, cap, count);
}

// host-only diagnostic: the small dense symmetric eigensolver used by the eigen / rank steps (no GPU needed)
#include "small_eig.h"
extern "C" int manisdp_test_sym_eig(const double* A, int32_t n, double* w, double* V) {
  if (!A || !w || n < 0) return MANISDP_E_ARG;
  std::vector<double> a(A, A + (size_t)n * n), ww, vv;
  if (!sym_eig(a, n, ww, vv, V != nullptr)) return MANISDP_E_NUMERIC;  // V == NULL: eigenvalues only
  for (int i = 0; i < n; ++i) w[i] = ww[i];
  for (size_t i = 0; V && i < (size_t)n * n; ++i) V[i] = vv[i];
  return MANISDP_OK;
}

// benchmark hook for the fused vector kernels (K5-K7): `reps` back-to-back launches timed with CUDA events.
//   which: 0 retract (24 np bytes), 1 tangent projection (24 np), 2 tCG update pass (56 np), 3 tCG direction pass (32 np)
// The tCG passes run on the solver's own workspace with a benign scalar state (alpha = beta = 1e-3, not stopped); the
// workspace is re-initialised by the next tr_solve.
__global__ void k_bench_state(RtrState* st) {
  st->stop = 0;
  st->branch = 0;
  st->alpha = 1e-3;
  st->beta = 1e-3;
  st->tau = 0.0;
  st->j = 0;
  st->maxinner = 1 << 30;
  st->mininner = 1 << 30;  // never satisfies the residual stop, so `stop` stays 0 across the repetitions
  st->model_value = 1e300;
  st->kappa = 0.1;
  st->theta = 1.0;
  st->norm_r0 = 1.0;
  st->z_r = 1.0;
  st->ticket = 0u;
}
extern "C" int manisdp_vec_bench(manisdp_t* h, int32_t which, int32_t reps, double* ms_per_launch, double* bytes) {
  if (!h || h->p <= 0 || reps < 1 || which < 0 || which > 3)
    return msdp_fail(h, MANISDP_E_ARG, "vec_bench: bad argument / no factor set");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const double np8 = 8.0 * (double)h->nloc * (double)h->ld;
  const double mult[4] = {3.0, 3.0, 7.0, 4.0};
  for (int pass = 0; pass < 2; ++pass) {  // pass 0 = warm-up
    k_bench_state<<<1, 1, 0, h->stream>>>(h->st);
    KERNEL_CHECK(h);
    if (pass == 1) CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    const int n_it = pass == 0 ? 2 : reps;
    for (int i = 0; i < n_it; ++i) {
      switch (which) {
        case 0: MSDP_TRY(msdp_launch_retract(h, h->Ybuf[h->pt], h->Uslot, h->Hslot, 0)); break;
        case 1: MSDP_TRY(msdp_launch_project(h, h->Ybuf[h->pt], h->Uslot, h->Hslot)); break;
        case 2: MSDP_TRY(msdp_launch_tcg_update(h, 0, 0, 1)); break;  // defer = 1: totals land in st->tmp only
        default: MSDP_TRY(msdp_launch_tcg_dir(h)); break;
      }
    }
    if (pass == 1) CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  }
  CUDA_TRY(h, cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (ms_per_launch) *ms_per_launch = (double)ms / reps;
  if (bytes) *bytes = mult[which] * np8;
  h->cache_valid = h->grad_valid = 0;
  return MANISDP_OK;
}
