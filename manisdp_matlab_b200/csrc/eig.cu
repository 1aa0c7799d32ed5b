// eig.cu -- the outer-loop pieces that touch n-sized data: KKT residues, the saddle-escape eigen step (K8: device
// LOBPCG on the dual slack S reusing the Hessian's SpMM / GEMM, replacing eig(full(S)) of
// ManiSDP_onlyunitdiag.m:50 / ManiSDP_unitdiag.m:68 / ManiSDP_unittrace.m:68 / ManiSDP.m:66), the rank step through
// the p x p Gram matrix (K9, replacing svd(Y), ManiSDP_unitdiag.m:72-74,93-96) and the escape / line-search update
// (ManiSDP_unitdiag.m:97-107,138-150).  Only O(k^2) numbers (Gram blocks, Ritz coefficients) visit the host.
#include <math.h>
#include <string.h>
#include <algorithm>
#include "affine.h"
#include "dist.h"
#include "gemm.h"
#include "kernels.cuh"
#include "rowops.cuh"
#include "small_eig.h"

#include <stdio.h>
#include <stdlib.h>
#include <chrono>
// coarse accounting of the eigen step (MANISDP_EIG_DEBUG=1 prints it at destroy): device wait vs host Rayleigh-Ritz
static double g_eig_wait_s = 0.0, g_eig_host_s = 0.0, g_eig_rr_s = 0.0;  // g_eig_rr_s: Rayleigh-Ritz arithmetic inside g_eig_host_s
static long g_eig_iters_total = 0;
// rank step (MANISDP_EIG_DEBUG): Gram kernel + copy, host eigenvalues, host eigenvectors, installing the cut factor
static double g_rank_gram_s = 0.0, g_rank_vals_s = 0.0, g_rank_vecs_s = 0.0, g_rank_install_s = 0.0;
static long g_rank_calls = 0;
struct ScopedSeconds {
  double& acc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit ScopedSeconds(double& a) : acc(a) {}
  ~ScopedSeconds() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

#define EIG_TR 32       // rows per shared-memory tile
#define EIG_MAXNB 48    // 3 * block size

// ---- block workspace ------------------------------------------------------------------------------------------------
struct EigWork {
  int k = 0, kld = 0;
  int64_t rows = 0;
  double *X = nullptr, *W = nullptr, *P = nullptr, *AX = nullptr, *AW = nullptr, *AP = nullptr;
  double* gather = nullptr;  // sharded: world*rows x kld
  double* gpart = nullptr;   // per-block partial Grams
  double* gout = nullptr;    // device: [G | GA | norms2]  (2*nb*nb + k)
  double* coef = nullptr;    // device: [Cx | Cw | Cp | theta]  (3*k*k + k)
  double* hbuf = nullptr;    // pinned mirror of gout
  double* hcoef = nullptr;   // pinned mirror of coef
  int nblocks = 0;
  int seg_chunks = 1;
};

static void eig_free(EigWork& w) {
  double* arrs[] = {w.X, w.W, w.P, w.AX, w.AW, w.AP, w.gather, w.gpart, w.gout, w.coef};
  for (double* a : arrs)
    if (a) cudaFree(a);
  if (w.hbuf) cudaFreeHost(w.hbuf);
  if (w.hcoef) cudaFreeHost(w.hcoef);
  w = EigWork();
}

static int eig_alloc(manisdp_handle* h, EigWork& w, int k) {
  const int kld = 4 * ((k + 3) / 4);
  const int64_t rows = msdp_rows_per_rank(h->n, h->world);
  if (w.k == k && w.rows == rows) return MANISDP_OK;
  eig_free(w);
  w.k = k;
  w.kld = kld;
  w.rows = rows;
  const size_t bytes = (size_t)rows * kld * sizeof(double);
  double** arrs[] = {&w.X, &w.W, &w.P, &w.AX, &w.AW, &w.AP};
  for (double** a : arrs) {
    CUDA_TRY(h, cudaMalloc((void**)a, bytes));
    CUDA_TRY(h, cudaMemset(*a, 0, bytes));
  }
  if (h->world > 1) {
    CUDA_TRY(h, cudaMalloc((void**)&w.gather, bytes * h->world));
    CUDA_TRY(h, cudaMemset(w.gather, 0, bytes * h->world));
  }
  w.nblocks = std::min<int64_t>(h->num_sms * 2, std::max<int64_t>(1, (h->nloc + EIG_TR - 1) / EIG_TR));
  const int nb = 3 * kld;
  const size_t gsz = (size_t)2 * nb * nb + kld;
  // wide blocks (nb > EIG_MAXNB): k_gram_seg writes `seg_chunks` partial [G | GA] pairs
  const int tiles = ((nb + 31) / 32) * ((nb + 31) / 32) * 2;
  w.seg_chunks = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)(4 * h->num_sms + tiles - 1) / tiles,
                                                                (h->nloc + EIG_TR - 1) / EIG_TR, (int64_t)64}));
  const size_t npart = std::max<size_t>((size_t)w.nblocks, (size_t)w.seg_chunks);
  CUDA_TRY(h, cudaMalloc((void**)&w.gpart, npart * gsz * sizeof(double)));
  CUDA_TRY(h, cudaMalloc((void**)&w.gout, gsz * sizeof(double)));
  CUDA_TRY(h, cudaMalloc((void**)&w.coef, ((size_t)3 * kld * kld + kld) * sizeof(double)));
  CUDA_TRY(h, cudaMallocHost((void**)&w.hbuf, gsz * sizeof(double)));
  CUDA_TRY(h, cudaMallocHost((void**)&w.hcoef, ((size_t)3 * kld * kld + kld) * sizeof(double)));
  return MANISDP_OK;
}

// ---- kernels --------------------------------------------------------------------------------------------------------
// G = S'S, GA = S'(AS) with S = [X | W | P] (n x 3*kld), plus (optionally) the squared column norms of W.
// Block partials -> last block sums them in block order (deterministic).
__global__ void __launch_bounds__(MSDP_THREADS)
    k_gram2(const double* X, const double* W, const double* P, const double* AX, const double* AW, const double* AP,
            int64_t nrows, int kld, double* gpart, double* gout, unsigned int* ticket) {
  extern __shared__ double smem[];
  const int nb = 3 * kld;
  double* s = smem;                 // EIG_TR x nb
  double* as = smem + EIG_TR * nb;  // EIG_TR x nb
  const int tid = threadIdx.x;
  const int nent = nb * nb;
  double accG[(EIG_MAXNB * EIG_MAXNB + MSDP_THREADS - 1) / MSDP_THREADS];
  double accA[(EIG_MAXNB * EIG_MAXNB + MSDP_THREADS - 1) / MSDP_THREADS];
  constexpr int EPT = (EIG_MAXNB * EIG_MAXNB + MSDP_THREADS - 1) / MSDP_THREADS;
#pragma unroll
  for (int e = 0; e < EPT; ++e) accG[e] = accA[e] = 0.0;
  for (int64_t r0 = (int64_t)blockIdx.x * EIG_TR; r0 < nrows; r0 += (int64_t)gridDim.x * EIG_TR) {
    const int tr = (int)min((int64_t)EIG_TR, nrows - r0);
    for (int i = tid; i < tr * nb; i += blockDim.x) {
      const int r = i / nb, c = i % nb, blk = c / kld, cc = c % kld;
      const size_t off = (size_t)(r0 + r) * kld + cc;
      const double* sp = blk == 0 ? X : (blk == 1 ? W : P);
      const double* ap = blk == 0 ? AX : (blk == 1 ? AW : AP);
      s[r * nb + c] = sp[off];
      as[r * nb + c] = ap[off];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = tid + e * MSDP_THREADS;
      if (idx < nent) {
        const int i = idx / nb, j = idx % nb;
        double g = 0.0, ga = 0.0;
        for (int r = 0; r < tr; ++r) {
          const double si = s[r * nb + i];
          g = fma(si, s[r * nb + j], g);
          ga = fma(si, as[r * nb + j], ga);
        }
        accG[e] += g;
        accA[e] += ga;
      }
    }
    __syncthreads();
  }
  const size_t gsz = (size_t)2 * nent;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int idx = tid + e * MSDP_THREADS;
    if (idx < nent) {
      gpart[(size_t)blockIdx.x * gsz + idx] = accG[e];
      gpart[(size_t)blockIdx.x * gsz + nent + idx] = accA[e];
    }
  }
  if (grid_last(ticket)) {
    __threadfence();
    for (int idx = tid; idx < 2 * nent; idx += blockDim.x) {
      double t = 0.0;
      for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(&gpart[(size_t)b * gsz + idx]);
      gout[idx] = t;
    }
  }
}

// [X, P, AX, AP] <- [X W P] * [Cx; Cw; Cp] , [W P] * [Cw; Cp] (and the same for the A-images), in place
__global__ void __launch_bounds__(MSDP_THREADS)
    k_combine(double* X, const double* W, double* P, double* AX, const double* AW, double* AP, int64_t nrows, int kld,
              const double* coef, int tile_rows) {
  extern __shared__ double smem[];
  double* cf = smem;                      // 3 * kld * kld
  double* tile = smem + 3 * kld * kld;    // 6 x tile_rows x kld
  const int tid = threadIdx.x;
  for (int i = tid; i < 3 * kld * kld; i += blockDim.x) cf[i] = coef[i];
  const int tsz = tile_rows * kld;
  for (int64_t r0 = (int64_t)blockIdx.x * tile_rows; r0 < nrows; r0 += (int64_t)gridDim.x * tile_rows) {
    const int tr = (int)min((int64_t)tile_rows, nrows - r0);
    __syncthreads();
    for (int i = tid; i < tr * kld; i += blockDim.x) {
      const size_t off = (size_t)r0 * kld + i;
      tile[0 * tsz + i] = X[off];
      tile[1 * tsz + i] = W[off];
      tile[2 * tsz + i] = P[off];
      tile[3 * tsz + i] = AX[off];
      tile[4 * tsz + i] = AW[off];
      tile[5 * tsz + i] = AP[off];
    }
    __syncthreads();
    for (int i = tid; i < tr * kld; i += blockDim.x) {
      const int r = i / kld, c = i % kld;
      double xs = 0.0, ps = 0.0, axs = 0.0, aps = 0.0;
      for (int q = 0; q < kld; ++q) {
        const double cx = cf[q * kld + c], cw = cf[kld * kld + q * kld + c], cp = cf[2 * kld * kld + q * kld + c];
        const double wv = tile[1 * tsz + r * kld + q], pv = tile[2 * tsz + r * kld + q];
        const double awv = tile[4 * tsz + r * kld + q], apv = tile[5 * tsz + r * kld + q];
        const double pp = cw * wv + cp * pv, app = cw * awv + cp * apv;
        ps += pp;
        aps += app;
        xs += cx * tile[0 * tsz + r * kld + q] + pp;
        axs += cx * tile[3 * tsz + r * kld + q] + app;
      }
      const size_t off = (size_t)r0 * kld + i;
      X[off] = xs;
      P[off] = ps;
      AX[off] = axs;
      AP[off] = aps;
    }
  }
}

// W = AX - X*diag(theta); gout[2*nb*nb + c] = |W(:,c)|^2
__global__ void __launch_bounds__(MSDP_THREADS)
    k_resid(const double* X, const double* AX, double* W, int64_t nrows, int kld, const double* theta, double* gpart,
            double* gout, unsigned int* ticket, int norm_off) {
  __shared__ double snorm[64];
  const int tid = threadIdx.x;
  if (tid < 64) snorm[tid] = 0.0;
  __syncthreads();
  // thread -> fixed column (blockDim % kld == 0 is guaranteed by the launcher)
  const int c = tid % kld;
  const double th = theta[c];
  double acc = 0.0;
  const int64_t total = nrows * kld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double w = AX[i] - th * X[i];
    W[i] = w;
    acc += w * w;
  }
  // deterministic in-block reduction per column: serialise through shared memory in thread order
  for (int pass = 0; pass < (int)blockDim.x / kld; ++pass) {
    if (tid / kld == pass) snorm[c] += acc;
    __syncthreads();
  }
  if (tid < kld) gpart[(size_t)blockIdx.x * kld + tid] = snorm[tid];
  if (grid_last(ticket)) {
    __threadfence();
    if (tid < kld) {
      double t = 0.0;
      for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(&gpart[(size_t)b * kld + tid]);
      gout[norm_off + tid] = t;
    }
  }
}

__global__ void k_fill_randn_block(double* X, int64_t nloc, int k, int kld, int64_t row0, uint64_t seed) {
  const int64_t total = nloc * kld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kld;
    const int c = (int)(i % kld);
    // cheap counter hash -> uniform in (-1, 1); quality is irrelevant for a starting block
    uint64_t x = (uint64_t)((row0 + r) * 64 + c) * 0x9E3779B97F4A7C15ull + seed * 0xBF58476D1CE4E5B9ull;
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    X[i] = (c < k) ? ((double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0) : 0.0;
  }
}

// general tall-skinny Gram: part[chunk] = A(rows of chunk)' * B(rows of chunk), A: n x lda (ka cols), B: n x ldb (kb)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_gram_general(const double* A, int lda, int ka, const double* B, int ldb, int kb, int64_t nrows, double* part) {
  __shared__ double sa[EIG_TR][33], sb[EIG_TR][33];
  const int ti = blockIdx.x, tj = blockIdx.y, chunk = blockIdx.z, nchunks = gridDim.z;
  const int tid = threadIdx.x;
  const int64_t rows_per = (nrows + nchunks - 1) / nchunks;
  const int64_t rbeg = rows_per * chunk, rend = min(nrows, rbeg + rows_per);
  double acc[4] = {0, 0, 0, 0};
  const int oi = tid / 32, oj = tid % 32;  // outputs (oi + 8*q, oj), q < 4
  for (int64_t r0 = rbeg; r0 < rend; r0 += EIG_TR) {
    const int tr = (int)min((int64_t)EIG_TR, rend - r0);
    for (int i = tid; i < EIG_TR * 32; i += blockDim.x) {
      const int r = i / 32, c = i % 32;
      const int ca = ti * 32 + c, cb = tj * 32 + c;
      sa[r][c] = (r < tr && ca < ka) ? A[(size_t)(r0 + r) * lda + ca] : 0.0;
      sb[r][c] = (r < tr && cb < kb) ? B[(size_t)(r0 + r) * ldb + cb] : 0.0;
    }
    __syncthreads();
    for (int r = 0; r < EIG_TR; ++r) {
      const double bv = sb[r][oj];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(sa[r][oi + 8 * q], bv, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ti * 32 + oi + 8 * q, j = tj * 32 + oj;
    if (i < ka && j < kb) part[(size_t)chunk * ka * kb + (size_t)i * kb + j] = acc[q];
  }
}

// Grams of a WIDE basis (3*kld > EIG_MAXNB, i.e. options.delta > 12): the register-accumulator kernel k_gram2 covers
// 48 x 48 entries only, so the same two matrices G = S'S and GA = S'(AS), S = [X | W | P], are produced tile by tile
// (32 x 32 outputs per block, row chunks over blockIdx.z, which == 0 -> G, 1 -> GA); k_sum_chunks adds the chunks in
// order (deterministic).  Column c of S lives in array c / kld at column c % kld.
struct Seg3 {
  const double* s[3];
  const double* as[3];
};
__global__ void __launch_bounds__(MSDP_THREADS)
    k_gram_seg(Seg3 q, int kld, int64_t nrows, int nchunks, double* part) {
  __shared__ double sa[EIG_TR][33], sb[EIG_TR][33];
  const int nb = 3 * kld;
  const int ti = blockIdx.x, tj = blockIdx.y;
  const int chunk = blockIdx.z % nchunks, which = blockIdx.z / nchunks;
  const int tid = threadIdx.x;
  const int64_t rows_per = (nrows + nchunks - 1) / nchunks;
  const int64_t rbeg = rows_per * chunk, rend = min(nrows, rbeg + rows_per);
  double acc[4] = {0, 0, 0, 0};
  const int oi = tid / 32, oj = tid % 32;
  for (int64_t r0 = rbeg; r0 < rend; r0 += EIG_TR) {
    const int tr = (int)min((int64_t)EIG_TR, rend - r0);
    for (int i = tid; i < EIG_TR * 32; i += blockDim.x) {
      const int r = i / 32, c = i % 32;
      const int ca = ti * 32 + c, cb = tj * 32 + c;
      double va = 0.0, vb = 0.0;
      if (r < tr && ca < nb) va = q.s[ca / kld][(size_t)(r0 + r) * kld + ca % kld];
      if (r < tr && cb < nb) vb = (which ? q.as[cb / kld] : q.s[cb / kld])[(size_t)(r0 + r) * kld + cb % kld];
      sa[r][c] = va;
      sb[r][c] = vb;
    }
    __syncthreads();
    for (int r = 0; r < EIG_TR; ++r) {
      const double bv = sb[r][oj];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fma(sa[r][oi + 8 * k], bv, acc[k]);
    }
    __syncthreads();
  }
  const size_t nent = (size_t)nb * nb;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = ti * 32 + oi + 8 * k, j = tj * 32 + oj;
    // layout: part[chunk][which][i][j] so that k_sum_chunks over 2*nent entries yields [G | GA]
    if (i < nb && j < nb) part[((size_t)chunk * 2 + which) * nent + (size_t)i * nb + j] = acc[k];
  }
}
__global__ void k_sum_chunks(const double* part, double* out, int64_t nent, int nchunks) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nent; i += (int64_t)gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int c = 0; c < nchunks; ++c) t += part[(size_t)c * nent + i];
    out[i] = t;
  }
}

// out (n x ldo, ko cols, zero padded) = A (n x lda, ka cols) * Cm (ka x ko, row-major) [+ B (n x ldb, kb cols) * Dm]
__global__ void __launch_bounds__(MSDP_THREADS)
    k_rows_times_small(const double* A, int lda, int ka, const double* Cm, const double* B, int ldb, int kb,
                       const double* Dm, double* out, int ldo, int ko, int64_t nrows) {
  const int64_t total = nrows * ldo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ldo;
    const int c = (int)(i % ldo);
    double s = 0.0;
    if (c < ko) {
      for (int q = 0; q < ka; ++q) s = fma(A[(size_t)r * lda + q], Cm[(size_t)q * ko + c], s);
      if (B)
        for (int q = 0; q < kb; ++q) s = fma(B[(size_t)r * ldb + q], Dm[(size_t)q * ko + c], s);
    }
    out[i] = s;
  }
}

// ---- operator S on a block -------------------------------------------------------------------------------------------
static int apply_S(manisdp_handle* h, EigWork& w, const double* V, double* AV, double sign) {
  (void)sign;
  if (h->kind == MANISDP_ONLYUNITDIAG) {
    const double* gather = V;
    if (h->world > 1) {  // all-gather of the thin block
      MSDP_TRY(msdp_dist_allgather_block(h, V, w.gather, msdp_rows_per_rank(h->n, h->world) * w.kld));
      gather = w.gather;
    }
    return msdp_spmm_shift(h, gather, V, AV, w.kld, h->zdiag);
  }
  return msdp_affine_apply_S(h, V, AV, w.kld);
}

// ---- LOBPCG ------------------------------------------------------------------------------------------------------------
// Smallest (want_largest = 0) or largest (= 1) `nwant` eigenpairs of S.  Block size k >= nwant.  Results: vals[0..k),
// w.X (n x kld) Ritz vectors, resid = max residual norm over the wanted pairs.
// Converged when the largest residual norm of the wanted pairs is <= max(tol_abs, tol_rel * max|theta|).
static int lobpcg(manisdp_handle* h, EigWork& w, int nwant, int want_largest, double tol_abs, double tol_rel, int maxit,
                  int warm, std::vector<double>& vals, double* resid_out, int* iters_out, bool* conv_out = nullptr) {
  NvtxRange nvtx_range(want_largest ? "manisdp:lobpcg(lambda_max)" : "manisdp:lobpcg(lambda_min block)");
  const int k = w.k, kld = w.kld, nb = 3 * kld, nent = nb * nb;
  const int64_t n = h->nloc;
  const double sgn = want_largest ? -1.0 : 1.0;
  cudaStream_t s = h->stream;
  const bool wide = nb > EIG_MAXNB;
  const size_t sm_gram = (size_t)2 * EIG_TR * nb * sizeof(double);
  // k_combine keeps the 3 coefficient blocks + a tile of the six arrays in shared memory: shrink the tile for wide blocks
  int tile_rows = EIG_TR;
  while (tile_rows > 2 && ((size_t)3 * kld * kld + 6 * (size_t)tile_rows * kld) * sizeof(double) > 200 * 1024) tile_rows /= 2;
  const size_t sm_comb = ((size_t)3 * kld * kld + 6 * (size_t)tile_rows * kld) * sizeof(double);
  // the attribute is per device and cheap to set: no process-wide "done" flag (handles may live on several GPUs)
  CUDA_TRY(h, cudaFuncSetAttribute(k_gram2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CUDA_TRY(h, cudaFuncSetAttribute(k_combine, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int rthreads = (MSDP_THREADS / kld) * kld;
  unsigned int* ticket = &h->st->ticket;
  if (!warm) {
    k_fill_randn_block<<<std::max(1, (int)std::min<int64_t>(h->num_sms * 4, (n * kld + 255) / 256)), 256, 0, s>>>(
        w.X, n, k, kld, h->row_begin, 0x5eed + (uint64_t)want_largest);
    KERNEL_CHECK(h);
  }
  CUDA_TRY(h, cudaMemsetAsync(w.P, 0, (size_t)w.rows * kld * sizeof(double), s));
  CUDA_TRY(h, cudaMemsetAsync(w.AP, 0, (size_t)w.rows * kld * sizeof(double), s));
  CUDA_TRY(h, cudaMemsetAsync(w.W, 0, (size_t)w.rows * kld * sizeof(double), s));
  CUDA_TRY(h, cudaMemsetAsync(w.AW, 0, (size_t)w.rows * kld * sizeof(double), s));
  MSDP_TRY(apply_S(h, w, w.X, w.AX, sgn));

  std::vector<double> G, GA, theta(kld, 0.0), norms(kld, 0.0);
  std::vector<std::vector<double>> hist;
  std::vector<int> act_w, act_p;  // active column indices
  bool haveW = false, haveP = false;
  double resid = INFINITY;
  int it = 0;
  bool converged = false;
  vals.assign(k, 0.0);
  for (it = 0; it <= maxit; ++it) {
    // Grams of the current basis (+ residual norms computed by the previous k_resid)
    if (!wide) {
      k_gram2<<<w.nblocks, MSDP_THREADS, sm_gram, s>>>(w.X, w.W, w.P, w.AX, w.AW, w.AP, n, kld, w.gpart, w.gout, ticket);
      KERNEL_CHECK(h);
    } else {
      Seg3 q{{w.X, w.W, w.P}, {w.AX, w.AW, w.AP}};
      dim3 grid((nb + 31) / 32, (nb + 31) / 32, 2 * w.seg_chunks);
      k_gram_seg<<<grid, MSDP_THREADS, 0, s>>>(q, kld, n, w.seg_chunks, w.gpart);
      KERNEL_CHECK(h);
      k_sum_chunks<<<std::max(1, (2 * nent + 255) / 256), 256, 0, s>>>(w.gpart, w.gout, (int64_t)2 * nent, w.seg_chunks);
      KERNEL_CHECK(h);
    }
    if (h->world > 1) MSDP_TRY(msdp_dist_allreduce_buf(h, w.gout, 2 * nent + kld));
    CUDA_TRY(h, cudaMemcpyAsync(w.hbuf, w.gout, ((size_t)2 * nent + kld) * sizeof(double), cudaMemcpyDeviceToHost, s));
    const auto t_wait0 = std::chrono::steady_clock::now();
    CUDA_TRY(h, cudaStreamSynchronize(s));
    const auto t_host0 = std::chrono::steady_clock::now();
    g_eig_wait_s += std::chrono::duration<double>(t_host0 - t_wait0).count();
    struct HostTimer {
      std::chrono::steady_clock::time_point t0;
      ~HostTimer() { g_eig_host_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    } host_timer{t_host0};
    g_eig_iters_total += 1;
    const double* hG = w.hbuf;
    const double* hGA = w.hbuf + nent;
    if (haveW) {
      resid = 0.0;
      act_w.clear();
      for (int c = 0; c < k; ++c) {
        norms[c] = sqrt(std::max(0.0, w.hbuf[2 * nent + c]));
        if (c < nwant) resid = std::max(resid, norms[c]);
        if (norms[c] > tol_abs * 0.1 && norms[c] == norms[c]) act_w.push_back(c);
      }
      if (resid != resid) return msdp_fail(h, MANISDP_E_NUMERIC, "lobpcg: NaN residual");
      double tmax = 0.0;
      for (int c = 0; c < nwant; ++c) tmax = std::max(tmax, fabs(vals[c]));
      const double tol_now = std::max(tol_abs, tol_rel * tmax);
      converged = resid <= tol_now;
      if (converged || it == maxit) break;
      if (act_w.empty()) break;
      // Ritz values converge quadratically in the residual: stop when the wanted ones have stopped moving
      // (change over the last 10 iterations below a tenth of the tolerance) although the residual test is not met
      hist.push_back(std::vector<double>(vals.begin(), vals.begin() + nwant));
      if (hist.size() > 10) {
        const std::vector<double>& old = hist[hist.size() - 11];
        double moved = 0.0;
        for (int c = 0; c < nwant; ++c) moved = std::max(moved, fabs(old[c] - vals[c]));
        if (moved <= 0.1 * tol_now) break;
      }
    }
    // basis index list: X (all k), active W, active P
    std::vector<int> idx;
    for (int c = 0; c < k; ++c) idx.push_back(c);
    if (haveW)
      for (int c : act_w) idx.push_back(kld + c);
    act_p.clear();
    if (haveP)
      for (int c : act_w) {
        if (hG[(size_t)(2 * kld + c) * nb + (2 * kld + c)] > 0.0) act_p.push_back(c);
      }
    // Rayleigh-Ritz on span[X, W_active, P_active] by canonical orthogonalisation: the unit-diagonal Gram matrix Gs is
    // eigen-decomposed and only directions with eigenvalue > 1e-10 * max are kept, so a nearly dependent W or P column
    // costs one direction instead of the whole block (a plain Cholesky here either fails or amplifies rounding).
    bool solved = false;
    std::vector<double> Cfull;  // nbasis x k coefficients w.r.t. the unscaled basis columns
    std::vector<double> ritz;
    {
      std::vector<int> bidx = idx;
      for (int c : act_p) bidx.push_back(2 * kld + c);
      const int m = (int)bidx.size();
      std::vector<double> dscale(m);
      for (int a = 0; a < m; ++a) {
        const double gaa = hG[(size_t)bidx[a] * nb + bidx[a]];
        dscale[a] = gaa > 0 ? 1.0 / sqrt(gaa) : 0.0;
      }
      std::vector<double> Gs((size_t)m * m), As((size_t)m * m);
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
          const double g = 0.5 * (hG[(size_t)bidx[a] * nb + bidx[b]] + hG[(size_t)bidx[b] * nb + bidx[a]]);
          const double ga = 0.5 * (hGA[(size_t)bidx[a] * nb + bidx[b]] + hGA[(size_t)bidx[b] * nb + bidx[a]]);
          Gs[(size_t)a * m + b] = g * dscale[a] * dscale[b];
          As[(size_t)a * m + b] = sgn * ga * dscale[a] * dscale[b];
        }
      // Fast path: Cholesky of the unit-diagonal Gram matrix (Gs = L L'), At = L^-1 As L^-T, ONE small eigen-solve instead of
      // the two of the canonical orthogonalisation (two 36 x 36 decompositions were ~80 % of the host time of an iteration:
      // 52 500 iterations on the q = 60 quartic, 9.8 s of its 22.6 s).
      {
        // Cholesky over the basis columns in order X | W | P; a W or P column whose squared pivot is below 1e-3 (less than
        // 3 % of it is outside the span of the columns before it) is DROPPED -- it gets zero coefficients, exactly what the
        // canonical orthogonalisation does to a dependent direction -- so that L stays well conditioned (the Ritz values
        // keep ~13 digits) and one eigen-solve per iteration suffices also on the degenerate spectra of theta SDPs, where
        // nearly every iteration has such a column.  A small pivot on an X column means the block itself degenerated: the
        // canonical path below takes over.
        std::vector<int> kept;
        std::vector<double> L((size_t)m * m, 0.0);  // row a of L (over the kept columns) at L[a*m ..]
        bool chol_ok = true;
        for (int a = 0; a < m && chol_ok; ++a) {
          const int nk = (int)kept.size();
          double* La = &L[(size_t)nk * m];  // provisional row: becomes row nk of the compact factor if the column is kept
          for (int b = 0; b < nk; ++b) {
            double sum = Gs[(size_t)a * m + kept[(size_t)b]];
            const double* Lb = &L[(size_t)b * m];
            for (int q = 0; q < b; ++q) sum -= La[q] * Lb[q];
            La[b] = sum / Lb[b];
          }
          double d = Gs[(size_t)a * m + a];
          for (int q = 0; q < nk; ++q) d -= La[q] * La[q];
          if (!(d > 1e-3)) {
            if (a < k) chol_ok = false;  // X column
            continue;                    // W / P column: dropped
          }
          La[nk] = sqrt(d);
          kept.push_back(a);
        }
        const int mk = (int)kept.size();
        if (chol_ok && mk >= k) {
          // T1 = L^-1 As(kept, kept)  (forward substitution, column by column), At = T1 L^-T = (L^-1 T1')'
          std::vector<double> T1((size_t)mk * mk), At((size_t)mk * mk);
          for (int c = 0; c < mk; ++c)
            for (int a = 0; a < mk; ++a) {
              double sum = As[(size_t)kept[(size_t)a] * m + kept[(size_t)c]];
              const double* La = &L[(size_t)a * m];
              for (int q = 0; q < a; ++q) sum -= La[q] * T1[(size_t)q * mk + c];
              T1[(size_t)a * mk + c] = sum / La[a];
            }
          for (int r = 0; r < mk; ++r)  // row r of T1: solve L x = T1(r, :)'  ->  At(r, :) = x'
            for (int a = 0; a < mk; ++a) {
              double sum = T1[(size_t)r * mk + a];
              const double* La = &L[(size_t)a * m];
              for (int q = 0; q < a; ++q) sum -= La[q] * At[(size_t)r * mk + q];
              At[(size_t)r * mk + a] = sum / La[a];
            }
          for (int r = 0; r < mk; ++r)
            for (int q = r + 1; q < mk; ++q) {
              const double v = 0.5 * (At[(size_t)r * mk + q] + At[(size_t)q * mk + r]);
              At[(size_t)r * mk + q] = At[(size_t)q * mk + r] = v;
            }
          std::vector<double> ev, Zt;
          bool okeig = sym_eig(At, mk, ev, Zt);
          for (int c = 0; okeig && c < k; ++c) okeig = std::isfinite(ev[c]);
          if (okeig) {
            // C(kept, :) = L^-T Zt(:, 0:k)  (back substitution), dropped columns keep zero coefficients; undo the scaling
            Cfull.assign((size_t)m * k, 0.0);
            std::vector<double> Ck((size_t)mk * k);
            for (int c = 0; c < k; ++c)
              for (int a = mk - 1; a >= 0; --a) {
                double sum = Zt[(size_t)a * mk + c];
                for (int q = a + 1; q < mk; ++q) sum -= L[(size_t)q * m + a] * Ck[(size_t)q * k + c];
                Ck[(size_t)a * k + c] = sum / L[(size_t)a * m + a];
              }
            for (int a = 0; a < mk; ++a)
              for (int c = 0; c < k; ++c) Cfull[(size_t)kept[(size_t)a] * k + c] = Ck[(size_t)a * k + c] * dscale[kept[(size_t)a]];
            ritz.assign(ev.begin(), ev.begin() + k);
            idx = bidx;
            solved = true;
          }
        }
      }
      std::vector<double> lam, Q;
      if (!solved && sym_eig(Gs, m, lam, Q) && lam[m - 1] > 0.0 && std::isfinite(lam[m - 1])) {
        std::vector<int> keep;
        for (int a = 0; a < m; ++a)
          if (lam[a] > 1e-10 * lam[m - 1]) keep.push_back(a);
        const int mk = (int)keep.size();
        if (mk >= k) {
          // Tm = Q(:, keep) * diag(lam_keep^-1/2)   (m x mk);  At = Tm' As Tm
          std::vector<double> Tm((size_t)m * mk), AT((size_t)m * mk), At((size_t)mk * mk);
          for (int a = 0; a < m; ++a)
            for (int q = 0; q < mk; ++q) Tm[(size_t)a * mk + q] = Q[(size_t)a * m + keep[q]] / sqrt(lam[keep[q]]);
          for (int a = 0; a < m; ++a)
            for (int q = 0; q < mk; ++q) {
              double s2 = 0.0;
              for (int b = 0; b < m; ++b) s2 += As[(size_t)a * m + b] * Tm[(size_t)b * mk + q];
              AT[(size_t)a * mk + q] = s2;
            }
          for (int r = 0; r < mk; ++r)
            for (int q = 0; q < mk; ++q) {
              double s2 = 0.0;
              for (int a = 0; a < m; ++a) s2 += Tm[(size_t)a * mk + r] * AT[(size_t)a * mk + q];
              At[(size_t)r * mk + q] = s2;
            }
          for (int r = 0; r < mk; ++r)
            for (int q = r + 1; q < mk; ++q) {
              const double v = 0.5 * (At[(size_t)r * mk + q] + At[(size_t)q * mk + r]);
              At[(size_t)r * mk + q] = At[(size_t)q * mk + r] = v;
            }
          std::vector<double> ev, Zt;
          bool okeig = sym_eig(At, mk, ev, Zt);
          for (int c = 0; okeig && c < k; ++c) okeig = std::isfinite(ev[c]);
          if (okeig) {
            Cfull.assign((size_t)m * k, 0.0);
            for (int a = 0; a < m; ++a)
              for (int c = 0; c < k; ++c) {
                double s2 = 0.0;
                for (int q = 0; q < mk; ++q) s2 += Tm[(size_t)a * mk + q] * Zt[(size_t)q * mk + c];
                Cfull[(size_t)a * k + c] = s2 * dscale[a];
              }
            ritz.assign(ev.begin(), ev.begin() + k);
            idx = bidx;
            solved = true;
          }
        }
      }
    }
    g_eig_rr_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count();
    if (!solved) {
      if (!haveW) return msdp_fail(h, MANISDP_E_NUMERIC, "lobpcg: starting block is rank deficient");
      break;  // basis degenerated: accept the current Ritz pairs
    }
    // scatter coefficients into Cx | Cw | Cp (kld x kld each) + theta
    double* hc = w.hcoef;
    memset(hc, 0, ((size_t)3 * kld * kld + kld) * sizeof(double));
    for (size_t a = 0; a < idx.size(); ++a) {
      const int blk = idx[a] / kld, col = idx[a] % kld;
      for (int c = 0; c < k; ++c) hc[(size_t)blk * kld * kld + (size_t)col * kld + c] = Cfull[a * k + c];
    }
    for (int c = 0; c < k; ++c) {
      theta[c] = sgn * ritz[c];
      hc[(size_t)3 * kld * kld + c] = theta[c];
      vals[c] = theta[c];
    }
    CUDA_TRY(h, cudaMemcpyAsync(w.coef, hc, ((size_t)3 * kld * kld + kld) * sizeof(double), cudaMemcpyHostToDevice, s));
    k_combine<<<w.nblocks, MSDP_THREADS, sm_comb, s>>>(w.X, w.W, w.P, w.AX, w.AW, w.AP, n, kld, w.coef, tile_rows);
    KERNEL_CHECK(h);
    haveP = haveW;
    // periodically refresh AX = S*X to stop drift of the implicitly updated images
    if (it > 0 && it % 25 == 0) MSDP_TRY(apply_S(h, w, w.X, w.AX, sgn));
    k_resid<<<w.nblocks, rthreads, 0, s>>>(w.X, w.AX, w.W, n, kld, w.coef + 3 * kld * kld, w.gpart, w.gout, ticket,
                                          2 * nent);
    KERNEL_CHECK(h);
    MSDP_TRY(apply_S(h, w, w.W, w.AW, sgn));
    haveW = true;
  }
  if (resid_out) *resid_out = resid;
  if (iters_out) *iters_out = it;
  if (conv_out) *conv_out = converged;
  return MANISDP_OK;
}

// dense fallback for tiny problems: S = apply_S(I), host eigen-decomposition
static int small_dense_eig(manisdp_handle* h, int delta, std::vector<double>& vals, double* lam_max) {
  const int n = (int)h->n;
  EigWork w;
  const int k = n;
  MSDP_TRY(eig_alloc(h, w, k));
  const int wld = w.kld;
  std::vector<double> I((size_t)n * wld, 0.0), S((size_t)n * wld);
  for (int i = 0; i < n; ++i) I[(size_t)i * wld + i] = 1.0;
  CUDA_TRY(h, cudaMemcpy(w.X, I.data(), I.size() * sizeof(double), cudaMemcpyHostToDevice));
  int rc = apply_S(h, w, w.X, w.AX, 1.0);
  if (rc == MANISDP_OK) {
    cudaError_t e = cudaMemcpyAsync(S.data(), w.AX, S.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = msdp_fail(h, MANISDP_E_CUDA, cudaGetErrorString(e));
  }
  eig_free(w);
  MSDP_TRY(rc);
  std::vector<double> A((size_t)n * n), ev, Z;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = 0.5 * (S[(size_t)i * wld + j] + S[(size_t)j * wld + i]);
  if (!sym_eig(A, n, ev, Z)) return msdp_fail(h, MANISDP_E_NUMERIC, "dense eig failed");
  const int kk = std::min(delta, n);
  const int kld = 4 * ((kk + 3) / 4);
  vals.assign(ev.begin(), ev.begin() + kk);
  *lam_max = ev[n - 1];
  std::vector<double> V((size_t)n * kld, 0.0);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < kk; ++c) V[(size_t)i * kld + c] = Z[(size_t)i * n + c];
  if (h->eigvecs) cudaFree(h->eigvecs);
  h->eigvecs = nullptr;
  CUDA_TRY(h, cudaMalloc((void**)&h->eigvecs, V.size() * sizeof(double)));
  CUDA_TRY(h, cudaMemcpy(h->eigvecs, V.data(), V.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->eig_k = kk;
  h->eig_kld = kld;
  return MANISDP_OK;
}

// ---- KKT --------------------------------------------------------------------------------------------------------------
struct EigStore {
  EigWork lo, hi;
  bool warm_lo = false, warm_hi = false;
  int64_t rank_version = -1;
  int rank_p = -1;
  std::vector<double> rank_ev, rank_Z, rank_G;
  bool rank_have_Z = false;
};
#include <map>
#include <mutex>
static std::map<manisdp_handle*, EigStore> g_store;
static std::mutex g_store_mu;
void msdp_eig_release(manisdp_handle* h) {
  std::lock_guard<std::mutex> lk(g_store_mu);
  if (getenv("MANISDP_EIG_DEBUG") && g_eig_iters_total > 0)
    fprintf(stderr, "[manisdp eig] iterations %ld, device wait %.3f s, host part (RR + launches) %.3f s of which Rayleigh-Ritz "
                    "arithmetic %.3f s\n", g_eig_iters_total, g_eig_wait_s, g_eig_host_s, g_eig_rr_s);
  if (getenv("MANISDP_EIG_DEBUG") && g_rank_calls > 0)
    fprintf(stderr, "[manisdp rank] %ld decompositions: gram+copy %.3f s, eigenvalues (incl. gram) %.3f s, eigenvectors %.3f s, "
                    "install %.3f s\n", g_rank_calls, g_rank_gram_s, g_rank_vals_s, g_rank_vecs_s, g_rank_install_s);
  auto it = g_store.find(h);
  if (it != g_store.end()) {
    eig_free(it->second.lo);
    eig_free(it->second.hi);
    g_store.erase(it);
  }
}

int msdp_kkt(manisdp_handle* h, int delta, double eig_tol, int update_dual, manisdp_kkt_info* out) {
  NvtxRange nvtx_range("manisdp:kkt+eig");
  memset(out, 0, sizeof(*out));
  if (delta < 1) delta = 1;
  // the LOBPCG block is delta + 4 columns; its kernels (k_resid, k_combine) are laid out for blocks of <= 64 columns
  if (delta > 60)
    return msdp_fail(h, MANISDP_E_ARG, "kkt: options.delta > 60 is not supported (LOBPCG block of delta + 4 <= 64 columns)");
  // eig_tol < 0: ADAPTIVE accuracy of the eigen step, |eig_tol| = the driver's KKT tolerance.  Far from convergence the
  // escape directions and dinf only need a few digits, so the residual tolerance follows the previous dinf
  // (1e-2 * dinf_prev, clamped to [1e-9, 1e-5]); whenever the resulting dinf comes within 100x of the KKT tolerance the
  // block is continued (warm) to the tight tolerance 1e-9 before anything is reported, so an accepted "dinf < tol" always
  // rests on the same accuracy as the fixed-tolerance mode.  (The reference's eig() is exact: ManiSDP_unitdiag.m:68.)
  double kkt_tol = 0.0;
  if (eig_tol < 0) {
    kkt_tol = -eig_tol;
    eig_tol = std::min(1e-5, std::max(1e-9, 1e-2 * h->last_dinf));
  }
  if (eig_tol == 0) eig_tol = 1e-9;
  // residues + dual slack operator of the driver
  if (h->kind == MANISDP_ONLYUNITDIAG) {
    // z = sum(C.*X) = eG of the cost kernel; obj = sum(z); S = C - diag(z)  (ManiSDP_onlyunitdiag.m:45-49)
    if (!(h->cache_valid && h->grad_valid)) {
      double f;
      MSDP_TRY(manisdp_cost(h, &f));
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->zdiag, h->eG[h->pt], (size_t)h->nloc * sizeof(double), cudaMemcpyDeviceToDevice,
                                h->stream));
    out->obj = 2.0 * h->st_host->fx;
    out->z_sum = out->obj;
    out->by = out->obj;
    out->pinf = 0.0;
    out->gap = 0.0;
  } else if (h->dual.on) {
    MSDP_TRY(msdp_dual_kkt(h, update_dual, out));  // ADMM step; the eigen step below works on X = eX - diag(z)
  } else {
    MSDP_TRY(msdp_affine_kkt(h, update_dual, out));
  }
  std::vector<double> vals;
  double lam_max = 0.0, resid = 0.0;
  int iters = 0, converged = 1;
  if (h->n <= 96 && h->world <= 1) {
    MSDP_TRY(small_dense_eig(h, delta, vals, &lam_max));
  } else {
    g_store_mu.lock();
    EigStore& st = g_store[h];
    g_store_mu.unlock();
    // lambda_max: small block, loose tolerance (only the normaliser 1 + lambda_max of dinf needs it)
    MSDP_TRY(eig_alloc(h, st.hi, 4));
    std::vector<double> hv;
    double r2 = 0.0;
    int it2 = 0;
    MSDP_TRY(lobpcg(h, st.hi, 1, 1, 0.0, 1e-6, 80, st.warm_hi ? 1 : 0, hv, &r2, &it2));
    st.warm_hi = true;
    lam_max = hv[0];
    const int k = 4 * ((delta + 4 + 3) / 4);
    const bool warm = st.warm_lo && st.lo.k == k;
    MSDP_TRY(eig_alloc(h, st.lo, k));
    const double tol_abs = eig_tol * (1.0 + fabs(lam_max));
    // Ritz values are upper bounds of lambda_min: an eigen step that stopped on stagnation / maxit without meeting
    // the residual test could report a dinf that is too small.  Continue from the current block (warm) up to twice
    // more, and tell the caller (eig_converged) if the residual test is still not met.
    bool conv = false;
    MSDP_TRY(lobpcg(h, st.lo, delta, 0, tol_abs, 0.0, 1500, warm ? 1 : 0, vals, &resid, &iters, &conv));
    double tol_now = tol_abs;
    if (kkt_tol > 0 && eig_tol > 1e-9 && !vals.empty() &&
        std::max(0.0, -vals[0]) / (1.0 + lam_max) < 100.0 * kkt_tol) {  // close to acceptance: tighten before reporting
      tol_now = 1e-9 * (1.0 + fabs(lam_max));
      int it3 = 0;
      MSDP_TRY(lobpcg(h, st.lo, delta, 0, tol_now, 0.0, 1500, 1, vals, &resid, &it3, &conv));
      iters += it3;
    }
    for (int attempt = 0; attempt < 2 && !conv; ++attempt) {
      int it3 = 0;
      MSDP_TRY(lobpcg(h, st.lo, delta, 0, tol_now, 0.0, 1500, 1, vals, &resid, &it3, &conv));
      iters += it3;
    }
    converged = conv ? 1 : 0;
    st.warm_lo = true;
    // keep the wanted vectors for manisdp_escape
    const int kld = st.lo.kld;
    if (h->eigvecs) cudaFree(h->eigvecs);
    h->eigvecs = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&h->eigvecs, (size_t)st.lo.rows * kld * sizeof(double)));
    CUDA_TRY(h, cudaMemcpyAsync(h->eigvecs, st.lo.X, (size_t)st.lo.rows * kld * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->eig_k = std::min(delta, st.lo.k);
    h->eig_kld = kld;
    vals.resize(h->eig_k);
  }
  if (h->eigvals_host) free(h->eigvals_host);
  h->eigvals_host = (double*)malloc(sizeof(double) * std::max<size_t>(1, vals.size()));
  int nneg = 0;
  for (size_t i = 0; i < vals.size(); ++i) {
    h->eigvals_host[i] = vals[i];
    if (vals[i] < 0.0) ++nneg;
  }
  out->lam_min = vals.empty() ? 0.0 : vals[0];
  out->lam_max = lam_max;
  out->dinf = std::max(0.0, -out->lam_min) / (1.0 + (h->dual.on ? fabs(lam_max) : lam_max));  // ManiSDP_onlyunitdiag.m:51 / ManiDSDP_unitdiag.m:87
  h->last_dinf = out->dinf;
  out->nneg = std::min(nneg, delta);
  out->eig_iters = iters;
  out->eig_resid = resid;
  out->eig_converged = converged;
  return MANISDP_OK;
}

// ---- rank step ----------------------------------------------------------------------------------------------------------
static int gram_general(manisdp_handle* h, const double* A, int lda, int ka, const double* B, int ldb, int kb,
                        std::vector<double>& out) {
  const int ti = (ka + 31) / 32, tj = (kb + 31) / 32;
  const size_t nent = (size_t)ka * kb;
  int nchunks = (int)std::min<int64_t>(std::max<int64_t>(1, (32ll << 20) / (int64_t)(nent * 8)), 148);
  nchunks = (int)std::min<int64_t>(nchunks, std::max<int64_t>(1, h->nloc / EIG_TR));
  double *part = nullptr, *dev = nullptr;
  MSDP_TRY(msdp_scratch(h, 1, nent * sizeof(double), (void**)&dev));
  cudaError_t e = cudaSuccess;
  if (ka >= 64 && kb >= 64) {
    // wide factors (BQP-60: p up to ~420): the Gram matrix is a K = n_local GEMM for the FP64 tensor pipe (split-K over
    // the rows, slices added in order) instead of 32 x 32 FMA tiles -- 9 ms -> well under 1 ms at p = 400
    MSDP_TRY(msdp_gemm_tn(h, A, lda, ka, B, ldb, kb, h->nloc, dev));
  } else if (e == cudaSuccess) {
    MSDP_TRY(msdp_scratch(h, 0, (size_t)nchunks * nent * sizeof(double), (void**)&part));
    dim3 grid(ti, tj, nchunks);
    k_gram_general<<<grid, MSDP_THREADS, 0, h->stream>>>(A, lda, ka, B, ldb, kb, h->nloc, part);
    k_sum_chunks<<<std::max(1, (int)((nent + 255) / 256)), 256, 0, h->stream>>>(part, dev, (int64_t)nent, nchunks);
    h->launches += 2;
    e = cudaPeekAtLastError();
  }
  int rc = MANISDP_OK;
  if (e == cudaSuccess && h->world > 1) rc = msdp_dist_allreduce_buf(h, dev, (int64_t)nent);
  out.resize(nent);
  if (e == cudaSuccess && rc == MANISDP_OK)
    e = cudaMemcpyAsync(out.data(), dev, nent * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (rc != MANISDP_OK) return rc;
  CUDA_TRY(h, e);
  return MANISDP_OK;
}

// new factor (n x pnew) = Y * Cm (+ V * Dm), installed as the current point with width pnew
static int install_combination(manisdp_handle* h, const std::vector<double>& Cm, int ka, const double* V, int ldv,
                               int kb, const std::vector<double>& Dm, int pnew) {
  const int64_t ldn = 4 * ((pnew + 3) / 4);
  const int64_t rows = msdp_rows_per_rank(h->n, h->world);
  double *tmp = nullptr, *dC = nullptr, *dD = nullptr;
  MSDP_TRY(msdp_scratch(h, 2, (size_t)rows * ldn * sizeof(double), (void**)&tmp));
  CUDA_TRY(h, cudaMemsetAsync(tmp, 0, (size_t)rows * ldn * sizeof(double), h->stream));
  {
    const size_t nc = std::max<size_t>(1, Cm.size()), nd = Dm.size();
    double* cd = nullptr;
    MSDP_TRY(msdp_scratch(h, 3, (nc + nd + 2) * sizeof(double), (void**)&cd));
    dC = cd;
    if (nd) dD = cd + nc;
  }
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dC, Cm.data(), Cm.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && dD)
    e = cudaMemcpyAsync(dD, Dm.data(), Dm.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && ka >= 64) {
    // wide factors: Y*Cm (+ V*Dm) through the FP64 tensor pipe; the padding columns of tmp stay zero (memset above)
    MSDP_TRY(msdp_gemm_rows_small(h, h->Ybuf[h->pt], (int)h->ld, ka, dC, pnew, pnew, h->nloc, tmp, (int)ldn, 1.0, 0.0));
    if (dD) MSDP_TRY(msdp_gemm_rows_small(h, V, ldv, kb, dD, pnew, pnew, h->nloc, tmp, (int)ldn, 1.0, 1.0));
    e = cudaStreamSynchronize(h->stream);
  } else if (e == cudaSuccess) {
    const int64_t total = h->nloc * ldn;
    k_rows_times_small<<<std::max(1, (int)std::min<int64_t>(h->num_sms * 8, (total + 255) / 256)), 256, 0, h->stream>>>(
        h->Ybuf[h->pt], (int)h->ld, ka, dC, dD ? V : nullptr, ldv, kb, dD, tmp, (int)ldn, pnew, h->nloc);
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
  }
  int rc = MANISDP_OK;
  if (e == cudaSuccess) {
    rc = msdp_resize(h, pnew);
    if (rc == MANISDP_OK) {
      e = cudaMemcpyAsync(h->Ybuf[h->pt], tmp, (size_t)rows * ldn * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
  }
  MSDP_TRY(rc);
  CUDA_TRY(h, e);
  return MANISDP_OK;
}

int msdp_rank_cut(manisdp_handle* h, double theta, int apply, int64_t* r_out, int64_t* pnew_out) {
  NvtxRange nvtx_range("manisdp:rank_cut");
  const int p = (int)h->p;
  // the drivers ask for the rank first and cut afterwards: keep the decomposition of the unchanged point
  g_store_mu.lock();
  EigStore& es = g_store[h];
  g_store_mu.unlock();
  if (es.rank_version != h->y_version || es.rank_p != p) {
    std::vector<double>& G = es.rank_G;
    g_rank_calls++;
    {
      ScopedSeconds t(g_rank_gram_s);
      MSDP_TRY(gram_general(h, h->Ybuf[h->pt], (int)h->ld, p, h->Ybuf[h->pt], (int)h->ld, p, G));
    }
    ScopedSeconds tv(g_rank_vals_s);
    for (int i = 0; i < p; ++i)
      for (int j = i + 1; j < p; ++j) {
        const double v = 0.5 * (G[(size_t)i * p + j] + G[(size_t)j * p + i]);
        G[(size_t)i * p + j] = G[(size_t)j * p + i] = v;
      }
    // the rank estimate needs the singular values only; eigenvectors are computed when a cut is actually applied
    std::vector<double> scratch;
    if (!sym_eig(G, p, es.rank_ev, scratch, false))
      return msdp_fail(h, MANISDP_E_NUMERIC, "rank_cut: eigen-decomposition failed");
    es.rank_have_Z = false;
    es.rank_version = h->y_version;
    es.rank_p = p;
  }
  const std::vector<double>& ev = es.rank_ev;
  {
    const double s1t = sqrt(std::max(0.0, ev[p - 1]));
    int rt = 0;
    for (int i = 0; i < p; ++i)
      if (h->rank_strict ? sqrt(std::max(0.0, ev[i])) > theta * s1t : sqrt(std::max(0.0, ev[i])) >= theta * s1t) ++rt;
    if (apply && rt <= p - 1 && rt >= 1 && !es.rank_have_Z) {
      std::vector<double> ev2;
      ScopedSeconds tz(g_rank_vecs_s);
      if (!sym_eig(es.rank_G, p, ev2, es.rank_Z, true))
        return msdp_fail(h, MANISDP_E_NUMERIC, "rank_cut: eigen-decomposition failed");
      es.rank_ev = ev2;
      es.rank_have_Z = true;
    }
  }
  const std::vector<double>& Z = es.rank_Z;
  // singular values of Y = sqrt(eigenvalues of Y'Y), descending (ManiSDP_unitdiag.m:72-74)
  const double s1 = sqrt(std::max(0.0, ev[p - 1]));
  int r = 0;
  for (int i = 0; i < p; ++i)
    if (h->rank_strict ? sqrt(std::max(0.0, ev[i])) > theta * s1 : sqrt(std::max(0.0, ev[i])) >= theta * s1) ++r;  // ManiDSDP_unitdiag.m:91 is strict
  if (r_out) *r_out = r;
  if (apply && r <= p - 1 && r >= 1) {
    ScopedSeconds ti(g_rank_install_s);
    // Y <- Y * U_r  (= V(:,1:r)'.*e(1:r) of the reference's p x n layout, :93-96)
    std::vector<double> Cm((size_t)p * r);
    for (int q = 0; q < p; ++q)
      for (int c = 0; c < r; ++c) Cm[(size_t)q * r + c] = Z[(size_t)q * p + (p - 1 - c)];
    MSDP_TRY(install_combination(h, Cm, p, nullptr, 0, 0, std::vector<double>(), r));
  }
  if (pnew_out) *pnew_out = h->p;
  return MANISDP_OK;
}

// ---- escape ---------------------------------------------------------------------------------------------------------------
int msdp_escape(manisdp_handle* h, int nne, double alpha, int line_search) {
  NvtxRange nvtx_range("manisdp:escape");
  if (nne < 0 || nne > h->eig_k) return msdp_fail(h, MANISDP_E_ARG, "escape: nne exceeds the eigenvectors kept by kkt");
  const int p = (int)h->p, pn = p + nne;
  // nne == 0 without line search: the reference still executes Y = [Y alpha*vS(:,1:0)]; Y = Y/norm(Y,'fro')
  // (ManiSDP_unittrace.m:111 / ManiSDP_unitdiag.m:106) -- after an applied rank cut the factor is off the sphere, so
  // the normalisation must not be skipped; only the Euclidean driver has nothing to do.
  if (nne == 0 && !line_search && h->mf == MF_EUCLID) return MANISDP_OK;
  // [Y, alpha*V] as one combination: Cm = [I_p 0], Dm = [0 alpha*I_nne]
  std::vector<double> Cm((size_t)p * pn, 0.0), Dm((size_t)std::max(1, nne) * pn, 0.0);
  for (int q = 0; q < p; ++q) Cm[(size_t)q * pn + q] = 1.0;
  const double a = line_search ? 0.0 : alpha;
  for (int q = 0; q < nne; ++q) Dm[(size_t)q * pn + p + q] = a;
  if (nne == 0) Dm.clear();
  MSDP_TRY(install_combination(h, Cm, p, h->eigvecs, h->eig_kld, nne, Dm, pn));
  if (line_search) {
    // stage U = [0, V] in SLOT_U for manisdp_line_search (the reference searches at the top of the NEXT outer
    // iteration, after sigma was updated: ManiSDP_unitdiag.m:54-56,108-112)
    const int64_t total = h->nloc * h->ld;
    std::vector<double> Z0((size_t)1, 0.0), D1((size_t)std::max(1, nne) * pn, 0.0);
    for (int q = 0; q < nne; ++q) D1[(size_t)q * pn + p + q] = 1.0;
    double *dz = nullptr, *dd = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&dd, D1.size() * sizeof(double)));
    CUDA_TRY(h, cudaMemcpy(dd, D1.data(), D1.size() * sizeof(double), cudaMemcpyHostToDevice));
    (void)dz;
    CUDA_TRY(h, cudaMemsetAsync(h->Uslot, 0, (size_t)msdp_rows_per_rank(h->n, h->world) * h->ld * sizeof(double),
                                h->stream));
    if (nne > 0) {
      k_rows_times_small<<<std::max(1, (int)std::min<int64_t>(h->num_sms * 8, (total + 255) / 256)), 256, 0,
                           h->stream>>>(h->eigvecs, h->eig_kld, nne, dd, nullptr, 0, 0, nullptr, h->Uslot, (int)h->ld,
                                        pn, h->nloc);
      h->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(dd);
    CUDA_TRY(h, e);
  } else if (h->mf != MF_EUCLID) {
    // Y = Y ./ sqrt(sum(Y.^2))  (:106) / Frobenius normalisation on the sphere
    CUDA_TRY(h, cudaMemsetAsync(h->eta[0], 0, (size_t)h->nloc * h->ld * sizeof(double), h->stream));
    MSDP_TRY(msdp_launch_retract(h, h->Ybuf[h->pt], h->eta[0], h->Ybuf[h->pt ^ 1], 0));
    CUDA_TRY(h, cudaMemcpyAsync(h->Ybuf[h->pt], h->Ybuf[h->pt ^ 1], (size_t)h->nloc * h->ld * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  h->cache_valid = h->grad_valid = 0;
  return MANISDP_OK;
}

// ---- line search along the staged escape direction (ManiSDP_unitdiag.m:138-150 and siblings) ---------------------------
__global__ void k_scaled_copy(const double* U, double a, double* out, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    double2 u = ld2(U + 2 * i);
    u.x *= a;
    u.y *= a;
    st2(out + 2 * i, u);
  }
}
__global__ void k_set_point(RtrState* st, int pt) { st->pt = pt; }

int msdp_line_search(manisdp_handle* h, double* alpha_out) {
  NvtxRange nvtx_range("manisdp:line_search");
  if (h->world > 1) return msdp_fail(h, MANISDP_E_ARG, "line search is not available on row-sharded handles");
  // co(Y): <C,YY'> for ONLYUNITDIAG (ManiSDP_onlyunitdiag.m:97-99), the AL cost otherwise
  const double scale = (h->kind == MANISDP_ONLYUNITDIAG) ? 2.0 : 1.0;
  h->cache_valid = 0;
  MSDP_TRY(msdp_ensure_costgrad(h));
  const double cost0 = scale * h->st_host->fx;
  const int64_t nvec = h->nloc * h->ld / 2;
  const int nb = std::max(1, (int)std::min<int64_t>(h->num_sms * 8, (nvec + 255) / 256));
  double alpha = 1.0;
  int i = 1;
  while (true) {
    k_scaled_copy<<<nb, 256, 0, h->stream>>>(h->Uslot, alpha, h->eta[0], nvec);
    KERNEL_CHECK(h);
    MSDP_TRY(msdp_launch_retract(h, h->Ybuf[h->pt], h->eta[0], h->Ybuf[h->pt ^ 1], 0));  // nY = normalise(Y + alpha*U)
    MSDP_TRY(msdp_costgrad(h, h->pt ^ 1, CG_COSTONLY));
    const int keep_pt = h->pt;
    MSDP_TRY(msdp_sync_state(h));
    h->pt = keep_pt;
    const double co = scale * h->st_host->tmp[0];
    if (i <= 15 && co - cost0 > -1e-3) {
      alpha *= 0.8;
      ++i;
      continue;
    }
    break;
  }
  // the candidate becomes the current point
  k_set_point<<<1, 1, 0, h->stream>>>(h->st, h->pt ^ 1);
  KERNEL_CHECK(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->pt ^= 1;
  h->y_version++;
  h->cache_valid = h->grad_valid = 0;
  if (alpha_out) *alpha_out = alpha;
  return MANISDP_OK;
}
