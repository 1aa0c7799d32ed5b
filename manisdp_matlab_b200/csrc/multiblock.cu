// multiblock.cu -- the multi-block driver's pieces (src/primal/ManiSDP_multiblock.m:60-160) on one device-resident factor.
//
// Layout.  The t blocks Y{1..t} (reference: p_i x n_i cells, unit columns on the first K.nob) are stacked into ONE array
// of N = sum(n_i) rows and ld = 4*ceil(max p_i / 4) columns; block i owns the rows [roff_i, roff_i + n_i) and the leading
// p_i columns, the other columns of its rows are zero.  Zero columns stay zero under 2*S*Y, 4*sigma*AyU*Y, the tangent
// projection, the retraction and every linear combination, so
//   * the manifold operations of multiblockmanifold.m:1-42 -- MEX files in the reference, src/C-files/innerc.cpp:20-32,
//     projc.cpp:19-56, retrc.cpp:24-47, lincombc.cpp:21-66 -- are the fused row kernels of tcg.cu with the per-row
//     manifold switch RtrState::nob_rows (rows of the unit-diagonal blocks come first), ONE launch for all blocks;
//   * the closures (:203-247) are the sparse-mode closures of affine.cu: At's rows are mapped to (row, column) of the
//     embedded block-diagonal matrix at create (affine.cu: `split`), the SDDMM / row-list kernels never see the block
//     boundaries, and nothing of size N x N is ever formed.
// What IS block-aware lives here: the per-block eigen-decomposition of the dual slack (:81-93), the per-block rank
// estimate / truncation (:116-129) and the per-block escape directions (:130-152), each batched over the blocks.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include "affine.h"
#include "kernels.cuh"
#include "rowops.cuh"
#include "small_eig.h"
#include <chrono>
#include <stdio.h>
#include <stdlib.h>

// per-phase wall seconds of the block-aware steps, printed at destroy when MANISDP_EIG_DEBUG is set
static double g_mb_resid_s = 0.0, g_mb_slack_s = 0.0, g_mb_eig_s = 0.0, g_mb_update_s = 0.0;
static long g_mb_kkt_calls = 0;
struct MbSeconds {
  double& acc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit MbSeconds(double& a) : acc(a) {}
  ~MbSeconds() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// ---- kernels ------------------------------------------------------------------------------------------------------------
// columns >= p[block(row)] <- 0
__global__ void __launch_bounds__(MSDP_THREADS)
    k_mb_mask(double* __restrict__ Y, const int* __restrict__ rowblk, const int* __restrict__ pw, int64_t nrows, int ld) {
  const int64_t total = nrows * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / ld;
    const int c = (int)(i - r * ld);
    if (c >= pw[rowblk[r]]) Y[i] = 0.0;
  }
}

// V (nrows x kld) <- columns [c0, c0 + kld) of the stacked identities: V[row, c] = (row - roff[block(row)] == c0 + c)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_mb_identity(double* __restrict__ V, const int* __restrict__ rowblk, const int* __restrict__ roff, int64_t nrows,
                  int kld, int c0) {
  const int64_t total = nrows * kld, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / kld;
    const int c = (int)(i - r * kld);
    V[i] = ((int)r - roff[rowblk[r]] == c0 + c) ? 1.0 : 0.0;
  }
}

// S{i} blocks out of the stacked product AV = S * I(:, c0 : c0 + kc): Sblk[off2[b] + a * n_b + c0 + c] = AV[roff[b] + a, c]
__global__ void __launch_bounds__(MSDP_THREADS)
    k_mb_pack_blocks(const double* __restrict__ AV, int kc, int c0, const int* __restrict__ rowblk,
                     const int* __restrict__ roff, const int64_t* __restrict__ off2, const int* __restrict__ nblk,
                     double* __restrict__ Sblk, int64_t nrows) {
  const int64_t total = nrows * kc, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / kc;
    const int c = (int)(i - r * kc);
    const int b = rowblk[r], nb = nblk[b];
    if (c0 + c < nb) Sblk[off2[b] + (int64_t)((int)r - roff[b]) * nb + c0 + c] = AV[i];
  }
}
// S{i} <- (S{i} + S{i}')/2, one CTA per block (the operator is symmetric up to the summation order of its rows)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_mb_symmetrize(double* __restrict__ Sblk, const int64_t* __restrict__ off2, const int* __restrict__ nblk) {
  const int b = blockIdx.x, nb = nblk[b];
  double* __restrict__ S = Sblk + off2[b];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e - i * nb;
    if (i < j) {
      const double v = 0.5 * (S[(size_t)i * nb + j] + S[(size_t)j * nb + i]);
      S[(size_t)i * nb + j] = v;
      S[(size_t)j * nb + i] = v;
    }
  }
}

// Gram matrices of all blocks, one CTA per block: G[blk][a][b] = sum_{rows of blk} Y[row, a] * Y[row, b]  (a, b < ld)
// (replaces svd(Y{i}), ManiSDP_multiblock.m:117: the singular values are the roots of the eigenvalues of Y_i' Y_i)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_mb_gram(const double* __restrict__ Y, int ld, const int* __restrict__ roff, double* __restrict__ G) {
  const int blk = blockIdx.x;
  const int r0 = roff[blk], r1 = roff[blk + 1];
  double* __restrict__ Gb = G + (size_t)blk * ld * ld;
  for (int e = threadIdx.x; e < ld * ld; e += blockDim.x) {
    const int a = e / ld, b = e - a * ld;
    double acc = 0.0;
    if (a <= b)
      for (int r = r0; r < r1; ++r) acc = fma(Y[(size_t)r * ld + a], Y[(size_t)r * ld + b], acc);
    Gb[e] = acc;  // upper triangle; the host mirrors it
  }
}

// new point of every block in one pass (ManiSDP_multiblock.m:121-128, 139-152):
//   out[row, c] = sum_q Y[row, q] * R[blk][q][c]     c <  rcut[blk]                 (rank-r truncation Y_i * W_r, or Y_i)
//               = a * V[row, c - rcut[blk]]           rcut[blk] <= c < rcut + nne    (escape directions; a = 0: line search)
//               = 0                                   otherwise
// and, when U != nullptr, the staged direction U[row, c] = V[row, c - rcut] on the escape columns, 0 elsewhere (:136-138)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_mb_recombine(const double* __restrict__ Y, int ldo, const double* __restrict__ R, const int* __restrict__ rowblk,
                   const int* __restrict__ pold, const int* __restrict__ rcut, const int* __restrict__ nne,
                   const double* __restrict__ evecs, const int64_t* __restrict__ off2, const int* __restrict__ roff,
                   const int* __restrict__ nblk, double a, double* __restrict__ out, double* __restrict__ U, int ldn,
                   int64_t nrows) {
  const int64_t total = nrows * ldn, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / ldn;
    const int c = (int)(i - r * ldn);
    const int blk = rowblk[r];
    const int rc = rcut[blk], ne = nne[blk];
    double v = 0.0, u = 0.0;
    if (c < rc) {
      const double* __restrict__ Rb = R + (size_t)blk * ldo * ldn;
      const int po = pold[blk];
      for (int q = 0; q < po; ++q) v = fma(Y[(size_t)r * ldo + q], Rb[(size_t)q * ldn + c], v);
    } else if (c < rc + ne) {
      u = evecs[off2[blk] + (int64_t)((int)r - roff[blk]) * nblk[blk] + (c - rc)];  // vS{i}(row, c - r_i)
      v = a * u;
    }
    out[i] = v;
    if (U) U[i] = u;
  }
}

// ---- host helpers -------------------------------------------------------------------------------------------------------
static int mb_grid(const manisdp_handle* h, int64_t total) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->num_sms * 8, (total + MSDP_THREADS - 1) / MSDP_THREADS));
}

static int mb_upload_widths(manisdp_handle* h) {
  std::vector<int> pw(h->mb_p.begin(), h->mb_p.end());
  CUDA_TRY(h, cudaMemcpyAsync(h->mb_pw, pw.data(), pw.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MANISDP_OK;
}

static int mb_check(const manisdp_handle* h) { return (h && h->kind == MANISDP_MULTIBLOCK) ? MANISDP_OK : MANISDP_E_ARG; }

// called from create before msdp_affine_setup: block tables + the per-row manifold switch
int msdp_mb_setup(manisdp_handle* h, const manisdp_problem* pb) {
  if (pb->nblocks < 1 || !pb->block_sizes) return msdp_fail(h, MANISDP_E_ARG, "MULTIBLOCK needs nblocks >= 1 and block_sizes");
  if (pb->nob < 0 || pb->nob > pb->nblocks) return msdp_fail(h, MANISDP_E_ARG, "MULTIBLOCK: nob must be in [0, nblocks]");
  const int t = pb->nblocks;
  h->mb_n.assign(pb->block_sizes, pb->block_sizes + t);
  h->mb_roff.assign((size_t)t + 1, 0);
  h->mb_off2.assign((size_t)t + 1, 0);
  for (int i = 0; i < t; ++i) {
    if (h->mb_n[i] < 1) return msdp_fail(h, MANISDP_E_ARG, "MULTIBLOCK: block orders must be >= 1");
    h->mb_roff[i + 1] = h->mb_roff[i] + h->mb_n[i];
    h->mb_off2[i + 1] = h->mb_off2[i] + h->mb_n[i] * h->mb_n[i];
  }
  if (h->mb_roff[t] != h->n) return msdp_fail(h, MANISDP_E_ARG, "MULTIBLOCK: n must equal sum(block_sizes)");
  h->mb_nob = pb->nob;
  h->mb_nob_rows = h->mb_roff[pb->nob];
  h->mb_p.assign((size_t)t, 0);
  std::vector<int> rowblk((size_t)h->n), roff(h->mb_roff.begin(), h->mb_roff.end());
  for (int i = 0; i < t; ++i)
    for (int64_t r = h->mb_roff[i]; r < h->mb_roff[i + 1]; ++r) rowblk[(size_t)r] = i;
  CUDA_TRY(h, cudaMalloc((void**)&h->mb_rowblk, rowblk.size() * sizeof(int)));
  CUDA_TRY(h, cudaMemcpy(h->mb_rowblk, rowblk.data(), rowblk.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMalloc((void**)&h->mb_roff_dev, roff.size() * sizeof(int)));
  CUDA_TRY(h, cudaMemcpy(h->mb_roff_dev, roff.data(), roff.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMalloc((void**)&h->mb_pw, (size_t)4 * t * sizeof(int)));  // widths + three work vectors of mb_update
  {  // stacked S{i} blocks (become the eigenvectors), Jacobi scratch, eigenvalues, their tables
    std::vector<int> nd(h->mb_n.begin(), h->mb_n.end());
    const size_t tot = (size_t)h->mb_off2[t];
    CUDA_TRY(h, cudaMalloc((void**)&h->mb_S, tot * sizeof(double)));
    CUDA_TRY(h, cudaMalloc((void**)&h->mb_V, tot * sizeof(double)));
    CUDA_TRY(h, cudaMalloc((void**)&h->mb_w, (size_t)h->n * sizeof(double)));
    CUDA_TRY(h, cudaMalloc((void**)&h->mb_off2_dev, ((size_t)t + 1) * sizeof(int64_t)));
    CUDA_TRY(h, cudaMemcpy(h->mb_off2_dev, h->mb_off2.data(), ((size_t)t + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMalloc((void**)&h->mb_n_dev, (size_t)t * sizeof(int)));
    CUDA_TRY(h, cudaMemcpy(h->mb_n_dev, nd.data(), (size_t)t * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMalloc((void**)&h->mb_sweeps, (size_t)t * sizeof(int)));
  }
  h->mb_eig_device = getenv("MANISDP_MB_EIG") && !strcmp(getenv("MANISDP_MB_EIG"), "device");
  const long long nob_rows = h->mb_nob_rows;
  CUDA_TRY(h, cudaMemcpy(&h->st->nob_rows, &nob_rows, sizeof(long long), cudaMemcpyHostToDevice));
  return MANISDP_OK;
}

void msdp_mb_free(manisdp_handle* h) {
  if (h->kind == MANISDP_MULTIBLOCK && g_mb_kkt_calls > 0 && getenv("MANISDP_EIG_DEBUG")) {
    fprintf(stderr, "[manisdp multiblock] %ld kkt steps: residues %.3f s, slack blocks (S*I + copy) %.3f s, eig of the blocks "
                    "%.3f s; rank cut + escape %.3f s\n", g_mb_kkt_calls, g_mb_resid_s, g_mb_slack_s, g_mb_eig_s, g_mb_update_s);
    g_mb_kkt_calls = 0;
    g_mb_resid_s = g_mb_slack_s = g_mb_eig_s = g_mb_update_s = 0.0;
  }
  if (h->mb_rowblk) cudaFree(h->mb_rowblk);
  if (h->mb_roff_dev) cudaFree(h->mb_roff_dev);
  if (h->mb_pw) cudaFree(h->mb_pw);
  void* more[] = {h->mb_S, h->mb_V, h->mb_w, h->mb_off2_dev, h->mb_n_dev, h->mb_sweeps};
  for (void* q : more)
    if (q) cudaFree(q);
  h->mb_S = h->mb_V = h->mb_w = nullptr;
  h->mb_off2_dev = nullptr;
  h->mb_n_dev = h->mb_sweeps = nullptr;
  h->mb_rowblk = h->mb_roff_dev = h->mb_pw = nullptr;
}

// M.typicaldist() and M.dim() of multiblockmanifold.m:3,11-15 at the current widths
double msdp_mb_typicaldist(const manisdp_handle* h) {
  double s = 0.0;
  for (size_t i = 0; i < h->mb_n.size(); ++i)
    s += ((int)i < h->mb_nob) ? M_PI * (double)h->mb_n[i] : (double)h->mb_p[i] * (double)h->mb_n[i];
  return sqrt(s);
}
double msdp_mb_dim(const manisdp_handle* h) {
  double s = 0.0;
  for (size_t i = 0; i < h->mb_n.size(); ++i)
    s += (double)(((int)i < h->mb_nob) ? h->mb_p[i] - 1 : h->mb_p[i]) * (double)h->mb_n[i];
  return s;
}

static int mb_set_widths(manisdp_handle* h, const int64_t* p, int64_t* pmax_out) {
  int64_t pmax = 0;
  for (size_t i = 0; i < h->mb_n.size(); ++i) {
    if (p[i] < 1) return msdp_fail(h, MANISDP_E_ARG, "multi-block widths must be >= 1");
    pmax = std::max(pmax, p[i]);
  }
  if (pmax > MSDP_MAX_LD) return msdp_fail(h, MANISDP_E_ARG, "multi-block widths must be <= 512");
  h->mb_p.assign(p, p + h->mb_n.size());
  *pmax_out = pmax;
  return MANISDP_OK;
}

// ---- state --------------------------------------------------------------------------------------------------------------
extern "C" int manisdp_mb_set_Y(manisdp_t* h, const double* Ycat, const int64_t* p) {
  if (mb_check(h) != MANISDP_OK || !Ycat || !p) return msdp_fail(h, MANISDP_E_ARG, "mb_set_Y: multi-block handle, Ycat and p needed");
  CUDA_TRY(h, cudaSetDevice(h->device));
  int64_t pmax = 0;
  MSDP_TRY(mb_set_widths(h, p, &pmax));
  std::vector<double> pad((size_t)h->n * pmax, 0.0);
  size_t src = 0;
  for (size_t i = 0; i < h->mb_n.size(); ++i)
    for (int64_t r = 0; r < h->mb_n[i]; ++r) {
      memcpy(&pad[(size_t)(h->mb_roff[i] + r) * pmax], Ycat + src, (size_t)p[i] * sizeof(double));
      src += (size_t)p[i];
    }
  MSDP_TRY(manisdp_set_Y(h, pad.data(), pmax, MANISDP_LAYOUT_ROWS));
  h->cache_valid = h->grad_valid = 0;
  h->y_version++;
  return mb_upload_widths(h);
}

extern "C" int manisdp_mb_get_Y(manisdp_t* h, double* Ycat) {
  if (mb_check(h) != MANISDP_OK || !Ycat) return msdp_fail(h, MANISDP_E_ARG, "mb_get_Y: multi-block handle and buffer needed");
  if (h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "no factor set");
  const int64_t pmax = h->p;
  std::vector<double> pad((size_t)h->n * pmax);
  MSDP_TRY(manisdp_get_Y(h, pad.data(), MANISDP_LAYOUT_ROWS));
  size_t dst = 0;
  for (size_t i = 0; i < h->mb_n.size(); ++i)
    for (int64_t r = 0; r < h->mb_n[i]; ++r) {
      memcpy(Ycat + dst, &pad[(size_t)(h->mb_roff[i] + r) * pmax], (size_t)h->mb_p[i] * sizeof(double));
      dst += (size_t)h->mb_p[i];
    }
  return MANISDP_OK;
}

extern "C" int manisdp_mb_get_widths(manisdp_t* h, int64_t* p) {
  if (mb_check(h) != MANISDP_OK || !p) return msdp_fail(h, MANISDP_E_ARG, "mb_get_widths: multi-block handle and buffer needed");
  std::copy(h->mb_p.begin(), h->mb_p.end(), p);
  return MANISDP_OK;
}

extern "C" int manisdp_mb_rand_Y(manisdp_t* h, const int64_t* p, uint64_t seed) {
  if (mb_check(h) != MANISDP_OK || !p) return msdp_fail(h, MANISDP_E_ARG, "mb_rand_Y: multi-block handle and widths needed");
  CUDA_TRY(h, cudaSetDevice(h->device));
  int64_t pmax = 0;
  MSDP_TRY(mb_set_widths(h, p, &pmax));
  MSDP_TRY(manisdp_rand_Y(h, pmax, seed));  // N(0,1) entries in all pmax columns (unit rows on the oblique blocks)
  MSDP_TRY(mb_upload_widths(h));
  // keep the leading p_i columns of block i and put the rows of the oblique blocks back on their spheres
  // (randc.cpp:52-80: a normalised Gaussian vector of p_i entries)
  double* Y = h->Ybuf[h->pt];
  k_mb_mask<<<mb_grid(h, h->n * h->ld), MSDP_THREADS, 0, h->stream>>>(Y, h->mb_rowblk, h->mb_pw, h->n, (int)h->ld);
  KERNEL_CHECK(h);
  CUDA_TRY(h, cudaMemsetAsync(h->eta[0], 0, (size_t)h->n * h->ld * sizeof(double), h->stream));
  MSDP_TRY(msdp_launch_retract(h, Y, h->eta[0], h->Ybuf[h->pt ^ 1], 0));
  CUDA_TRY(h, cudaMemcpyAsync(Y, h->Ybuf[h->pt ^ 1], (size_t)h->n * h->ld * sizeof(double), cudaMemcpyDeviceToDevice,
                              h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->cache_valid = h->grad_valid = 0;
  h->y_version++;
  return MANISDP_OK;
}

// ---- KKT step (ManiSDP_multiblock.m:66-97) ----------------------------------------------------------------------------
extern "C" int manisdp_mb_kkt(manisdp_t* h, int32_t update_dual, manisdp_kkt_info* out, double* dinfs, int32_t* nneg) {
  if (mb_check(h) != MANISDP_OK || !out) return msdp_fail(h, MANISDP_E_ARG, "mb_kkt: multi-block handle and out needed");
  if (h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "mb_kkt: no factor set");
  NvtxRange nvtx_range("manisdp:mb_kkt");
  CUDA_TRY(h, cudaSetDevice(h->device));
  memset(out, 0, sizeof(*out));
  // obj, pinf, y <- y - sigma*Axb, by = b'y + sum of z over the unit-diagonal blocks (:73-79, :84-88); zdiag holds z
  g_mb_kkt_calls++;
  {
    MbSeconds t(g_mb_resid_s);
    MSDP_TRY(msdp_affine_kkt(h, update_dual, out));
  }
  // S{i} of every block: S * (stacked identity) -- S is block diagonal, so the rows of block i hold S{i} (:81-89)
  const int t = (int)h->mb_n.size();
  int64_t nmax = 0;
  for (int64_t v : h->mb_n) nmax = std::max(nmax, v);
  const int kc = (int)std::min<int64_t>(4 * ((nmax + 3) / 4), MSDP_MAX_LD);  // identity columns per pass
  double *V = nullptr, *AV = nullptr;
  MSDP_TRY(msdp_scratch(h, 0, (size_t)h->n * kc * sizeof(double), (void**)&V));
  MSDP_TRY(msdp_scratch(h, 1, (size_t)h->n * kc * sizeof(double), (void**)&AV));
  const auto t_slack0 = std::chrono::steady_clock::now();
  for (int64_t c0 = 0; c0 < nmax; c0 += kc) {
    k_mb_identity<<<mb_grid(h, h->n * kc), MSDP_THREADS, 0, h->stream>>>(V, h->mb_rowblk, h->mb_roff_dev, h->n, kc, (int)c0);
    KERNEL_CHECK(h);
    MSDP_TRY(msdp_affine_apply_S(h, V, AV, kc));
    k_mb_pack_blocks<<<mb_grid(h, h->n * kc), MSDP_THREADS, 0, h->stream>>>(AV, kc, (int)c0, h->mb_rowblk, h->mb_roff_dev,
                                                                          h->mb_off2_dev, h->mb_n_dev, h->mb_S, h->n);
    KERNEL_CHECK(h);
  }
  k_mb_symmetrize<<<t, MSDP_THREADS, 0, h->stream>>>(h->mb_S, h->mb_off2_dev, h->mb_n_dev);
  KERNEL_CHECK(h);
  // eig(S{i}, 'vector') of every block (:90).  Default: Householder + QL per block on host threads (small_eig.h), the
  // eigenvectors uploaded next to the blocks.  MANISDP_MB_EIG=device (read at create): batched Jacobi, one CTA per block,
  // the blocks never leave the GPU (jacobi.cu) -- measured SLOWER on one B200 + 16 host threads (20 blocks of order 211:
  // 148 ms against 17 ms per KKT step; 89 blocks of order <= 55: 2.3 against 1.1 ms): a cyclic Jacobi needs ~10 sweeps of
  // 3 n^3 flops in ~2000 latency-bound rounds per block, QL ~9 n^3 in cache.  Kept as a tested alternative.
  h->mb_evals.assign((size_t)h->n, 0.0);
  std::vector<int> ok((size_t)t, 1);
  if (nmax <= 1024 && h->mb_eig_device) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    g_mb_slack_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_slack0).count();
    MbSeconds teig(g_mb_eig_s);
    MSDP_TRY(msdp_jacobi_batched(h, h->mb_S, h->mb_V, h->mb_w, h->mb_off2_dev, h->mb_roff_dev, h->mb_n_dev, h->mb_sweeps, t,
                                 (int)nmax));
    std::vector<int> sweeps((size_t)t);
    CUDA_TRY(h, cudaMemcpyAsync(h->mb_evals.data(), h->mb_w, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(sweeps.data(), h->mb_sweeps, (size_t)t * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < t; ++i) ok[(size_t)i] = sweeps[(size_t)i] > 0 || h->mb_n[i] <= 1;
  } else {
    std::vector<double> Sb((size_t)h->mb_off2[t]);  // S{i}, row-major n_i x n_i, at mb_off2[i]
    CUDA_TRY(h, cudaMemcpyAsync(Sb.data(), h->mb_S, Sb.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    g_mb_slack_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_slack0).count();
    MbSeconds teig(g_mb_eig_s);
    std::vector<double> Zall((size_t)h->mb_off2[t], 0.0);
    auto work = [&](int i) {
      const int ni = (int)h->mb_n[i];
      std::vector<double> A(Sb.begin() + h->mb_off2[i], Sb.begin() + h->mb_off2[i + 1]), ev, Z;
      if (!sym_eig(A, ni, ev, Z, true, t >= 4 ? 1 : 0)) {  // many blocks: one decomposition per thread, no nested threads
        ok[(size_t)i] = 0;
        return;
      }
      std::copy(ev.begin(), ev.end(), h->mb_evals.begin() + h->mb_roff[i]);
      std::copy(Z.begin(), Z.end(), Zall.begin() + h->mb_off2[i]);
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int T = (int)std::min<unsigned>({hw, 16u, (unsigned)t});
    if (T <= 1) {
      for (int i = 0; i < t; ++i) work(i);
    } else {
      // largest blocks first, round-robin over the threads
      std::vector<int> order((size_t)t);
      for (int i = 0; i < t; ++i) order[(size_t)i] = i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->mb_n[a] > h->mb_n[b]; });
      std::vector<std::thread> th;
      for (int w = 0; w < T; ++w)
        th.emplace_back([&, w]() {
          for (int q = w; q < t; q += T) work(order[(size_t)q]);
        });
      for (auto& x : th) x.join();
    }
    // the eigenvectors go where the device path leaves them
    CUDA_TRY(h, cudaMemcpy(h->mb_S, Zall.data(), Zall.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < t; ++i)
    if (!ok[(size_t)i]) return msdp_fail(h, MANISDP_E_NUMERIC, "mb_kkt: eigen-decomposition of a block failed");
  double dinf = 0.0, lmin = INFINITY, lmax = -INFINITY;
  int nneg_max = 0;
  for (int i = 0; i < t; ++i) {
    const double* ev = &h->mb_evals[(size_t)h->mb_roff[i]];
    const int ni = (int)h->mb_n[i];
    const double di = std::max(0.0, -ev[0]) / (1.0 + fabs(ev[ni - 1]));  // :91
    int cnt = 0;
    for (int a = 0; a < ni; ++a) cnt += (ev[a] < 0.0);
    if (dinfs) dinfs[i] = di;
    if (nneg) nneg[i] = cnt;
    dinf = std::max(dinf, di);
    lmin = std::min(lmin, ev[0]);
    lmax = std::max(lmax, ev[ni - 1]);
    nneg_max = std::max(nneg_max, cnt);
  }
  out->dinf = dinf;  // :93
  out->lam_min = lmin;
  out->lam_max = lmax;
  out->nneg = nneg_max;
  out->eig_iters = 0;
  out->eig_resid = 0.0;
  out->eig_converged = 1;
  h->last_dinf = dinf;
  h->mb_have_eigs = 1;
  return MANISDP_OK;
}

extern "C" int manisdp_mb_get_block_eigs(manisdp_t* h, int32_t blk, double* vals, double* vecs) {
  if (mb_check(h) != MANISDP_OK) return msdp_fail(h, MANISDP_E_ARG, "mb_get_block_eigs: multi-block handle needed");
  if (!h->mb_have_eigs) return msdp_fail(h, MANISDP_E_STATE, "mb_get_block_eigs: call manisdp_mb_kkt first");
  if (blk < 0 || blk >= (int)h->mb_n.size()) return msdp_fail(h, MANISDP_E_ARG, "mb_get_block_eigs: bad block index");
  const size_t ni = (size_t)h->mb_n[blk];
  if (vals) std::copy_n(h->mb_evals.begin() + h->mb_roff[blk], ni, vals);
  if (vecs) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpy(vecs, h->mb_S + h->mb_off2[blk], ni * ni * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return MANISDP_OK;
}

// ---- rank cut + escape of every block (ManiSDP_multiblock.m:114-153) --------------------------------------------------
extern "C" int manisdp_mb_update(manisdp_t* h, double theta, int32_t delta, double alpha, int32_t line_search,
                                 int32_t min_facsize, int64_t* p_new) {
  if (mb_check(h) != MANISDP_OK) return msdp_fail(h, MANISDP_E_ARG, "mb_update: multi-block handle needed");
  if (!h->mb_have_eigs) return msdp_fail(h, MANISDP_E_STATE, "mb_update: call manisdp_mb_kkt first");
  if (delta < 0) return msdp_fail(h, MANISDP_E_ARG, "mb_update: delta must be >= 0");
  NvtxRange nvtx_range("manisdp:mb_update");
  MbSeconds tupd(g_mb_update_s);
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int t = (int)h->mb_n.size();
  const int ldo = (int)h->ld;
  const double* Y = h->Ybuf[h->pt];
  // Gram matrices of all blocks in one launch, decomposed on the host (p_i x p_i, a few dozen entries each)
  double* Gd = nullptr;
  MSDP_TRY(msdp_scratch(h, 0, (size_t)t * ldo * ldo * sizeof(double), (void**)&Gd));
  k_mb_gram<<<t, MSDP_THREADS, 0, h->stream>>>(Y, ldo, h->mb_roff_dev, Gd);
  KERNEL_CHECK(h);
  std::vector<double> G((size_t)t * ldo * ldo);
  CUDA_TRY(h, cudaMemcpyAsync(G.data(), Gd, G.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  std::vector<int> pold(h->mb_p.begin(), h->mb_p.end()), rcut((size_t)t), nne((size_t)t, 0);
  std::vector<std::vector<double>> W((size_t)t);  // truncation basis of the blocks that are cut: p_i x r_i, row-major
  std::vector<int> okb((size_t)t, 1);
  auto block_rules = [&](int i) {
    const int pi = pold[(size_t)i], ni = (int)h->mb_n[i];
    rcut[(size_t)i] = pi;
    if (ni < min_facsize) return;  // :115
    if (pi > 1) {                  // :116-129
      std::vector<double> A((size_t)pi * pi), ev, Z;
      const double* Gb = &G[(size_t)i * ldo * ldo];
      for (int a = 0; a < pi; ++a)
        for (int b = a; b < pi; ++b) A[(size_t)a * pi + b] = A[(size_t)b * pi + a] = Gb[(size_t)a * ldo + b];
      // eigenvalues first (cheap); the basis only when the block is actually cut
      if (!sym_eig(A, pi, ev, Z, false, 1)) {
        okb[(size_t)i] = 0;
        return;
      }
      const double s1 = sqrt(std::max(0.0, ev[(size_t)pi - 1]));
      int r = 0;
      for (int a = 0; a < pi; ++a) r += (sqrt(std::max(0.0, ev[(size_t)a])) >= theta * s1);  // :123
      if (r == 0) r = 1;                                                                      // :124-126
      if (r < pi) {  // :127-130: Y{i} = diag(e(1:r))*V(:,1:r)'  ==  (row layout) Y_i * W(:, 1:r), singular values descending
        if (!sym_eig(A, pi, ev, Z, true, 1)) {
          okb[(size_t)i] = 0;
          return;
        }
        W[(size_t)i].assign((size_t)pi * r, 0.0);
        for (int q = 0; q < pi; ++q)
          for (int c = 0; c < r; ++c) W[(size_t)i][(size_t)q * r + c] = Z[(size_t)q * pi + (pi - 1 - c)];
        rcut[(size_t)i] = r;
      }
    }
    const double* ev = &h->mb_evals[(size_t)h->mb_roff[i]];
    int cnt = 0;
    for (int a = 0; a < ni; ++a) cnt += (ev[a] < 0.0);
    int ne = std::min(cnt, (int)delta);
    if (i < h->mb_nob) ne = std::max(ne, 1);        // :131-135
    if (rcut[(size_t)i] + ne > ni) ne = 0;          // :136-138
    nne[(size_t)i] = ne;
  };
  {  // the blocks are independent: spread them over host threads
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int T = (int)std::min<unsigned>({hw, 16u, (unsigned)t});
    if (T <= 1) {
      for (int i = 0; i < t; ++i) block_rules(i);
    } else {
      std::vector<std::thread> th;
      for (int w = 0; w < T; ++w)
        th.emplace_back([&, w]() {
          for (int i = w; i < t; i += T) block_rules(i);
        });
      for (auto& x : th) x.join();
    }
  }
  for (int i = 0; i < t; ++i)
    if (!okb[(size_t)i]) return msdp_fail(h, MANISDP_E_NUMERIC, "mb_update: Gram eigen-decomposition failed");
  int64_t pmax = 1;
  std::vector<int64_t> pn((size_t)t);
  for (int i = 0; i < t; ++i) {
    pn[(size_t)i] = rcut[(size_t)i] + nne[(size_t)i];
    pmax = std::max(pmax, pn[(size_t)i]);
  }
  if (pmax > MSDP_MAX_LD) return msdp_fail(h, MANISDP_E_ARG, "mb_update: factor width would exceed 512");
  const int ldn = (int)(4 * ((pmax + 3) / 4));
  // per-block combination matrices R (ldo x ldn): the truncation basis, or the identity on the kept columns
  std::vector<double> R((size_t)t * ldo * ldn, 0.0);
  for (int i = 0; i < t; ++i) {
    double* Rb = &R[(size_t)i * ldo * ldn];
    const int r = rcut[(size_t)i];
    if (!W[(size_t)i].empty()) {
      for (int q = 0; q < pold[(size_t)i]; ++q)
        for (int c = 0; c < r; ++c) Rb[(size_t)q * ldn + c] = W[(size_t)i][(size_t)q * r + c];
    } else {
      for (int q = 0; q < r; ++q) Rb[(size_t)q * ldn + q] = 1.0;
    }
  }
  // escape directions: the nne_i lowest eigenvectors of S{i}, read by the kernel from the stacked blocks mb_kkt left on
  // the device (h->mb_S)
  double *Rd = nullptr, *tmp = nullptr, *Utmp = nullptr;
  MSDP_TRY(msdp_scratch(h, 1, R.size() * sizeof(double), (void**)&Rd));
  const size_t out_elems = (size_t)h->n * ldn;
  MSDP_TRY(msdp_scratch(h, 2, out_elems * (line_search ? 2 : 1) * sizeof(double), (void**)&tmp));
  if (line_search) Utmp = tmp + out_elems;
  int* iw = h->mb_pw + t;  // device work vectors: pold | rcut | nne
  std::vector<int> pack;
  pack.insert(pack.end(), pold.begin(), pold.end());
  pack.insert(pack.end(), rcut.begin(), rcut.end());
  pack.insert(pack.end(), nne.begin(), nne.end());
  CUDA_TRY(h, cudaMemcpyAsync(iw, pack.data(), pack.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(Rd, R.data(), R.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  k_mb_recombine<<<mb_grid(h, (int64_t)out_elems), MSDP_THREADS, 0, h->stream>>>(
      Y, ldo, Rd, h->mb_rowblk, iw, iw + t, iw + 2 * t, h->mb_S, h->mb_off2_dev, h->mb_roff_dev, h->mb_n_dev,
      line_search ? 0.0 : alpha, tmp, Utmp, ldn, h->n);
  KERNEL_CHECK(h);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  // install: the arrays are re-laid out for the new width, the new point goes in as the current one
  MSDP_TRY(msdp_resize(h, pmax));
  h->mb_p = pn;
  MSDP_TRY(mb_upload_widths(h));
  double* Yn = h->Ybuf[h->pt];
  if (line_search) {
    // Y{i} = [Y{i}; zeros], U{i} = [zeros; vS'] staged for manisdp_line_search at the top of the next iteration (:62-64,:139-147)
    CUDA_TRY(h, cudaMemcpyAsync(Yn, tmp, out_elems * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->Uslot, Utmp, out_elems * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  } else {
    // Y{i} = [Y{i}; alpha*vS'] and, on the unit-diagonal blocks, Y{i}./sqrt(sum(Y{i}.^2)) (:148-152) == retr(., 0)
    CUDA_TRY(h, cudaMemsetAsync(h->eta[0], 0, out_elems * sizeof(double), h->stream));
    MSDP_TRY(msdp_launch_retract(h, tmp, h->eta[0], Yn, 0));
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->cache_valid = h->grad_valid = 0;
  h->y_version++;
  h->mb_have_eigs = 0;
  if (p_new) std::copy(pn.begin(), pn.end(), p_new);
  return MANISDP_OK;
}
