// dual.cu -- the dual approach: Riemannian ADMM on the SOS form of the relaxation (src/dual/ManiDSDP_unitdiag.m:28-194).
//
//   sup <C, X> + <c_f, w>   s.t.  A(X) + B(w) = b,  X >= 0,   dual slack S = Y Y' with unit diagonal (rows of Y on spheres)
//
// Reference closures (:171-191), with iA' = D^-1 A, D = diag(A A') (`options.dAAt`), P = A' D^-1 A:
//   sc = vec(S) - c,  y = D^-1 A sc,  As = P sc - sc - x/sigma,  Af = B'y - c_f - w/sigma
//   f  = b'y + sigma/2 (|As|^2 + |Af|^2),   X = mat(bA - sigma*As),   eG = 2 X Y,
//   eH = 2 X U - 4 sigma P(Y U') Y + 2 sigma (Y (U'Y) + U (Y'Y))
//
// How it maps onto the engine.  With A~ = D^-1/2 A (so P = A~' A~; rows of a SOS constraint matrix partition the
// positions of S, hence A A' is diagonal and P is an orthogonal projector with P x = 0 for every ADMM multiplier x the
// iteration produces, :78),
//   f(Y) = <C_eff, S> + sigma/2 |S|_F^2 - sigma/2 |A~ vec(S) - A~ c|^2 + sigma/2 |Af|^2 + k0
//   C_eff = bA + x - sigma*C,     k0 = -<bA + x, C> + sigma/2 |C|^2 + |x|^2 / (2 sigma)
//   X     = [C_eff - sigma * A~'(A~ vec(S) - A~ c)] + sigma*S
// i.e. the closures of the primal unit-diagonal driver (affine.cu: SDDMM / gather, row-list / DMMA GEMM) evaluated with
// the data (C_eff, A~, b_eff = A~ c, y = 0) and the penalty -sigma, plus three terms that only need the small Gram
// matrices Y'Y and Y'U (p x p): sigma/2 |Y'Y|_F^2 in the cost, 2 sigma Y (Y'Y) in the gradient and
// 2 sigma (Y (Y'U + U'Y) + U (Y'Y)) in the Hessian.  S = Y Y' is never formed inside the trust-region loop; the kernels
// below add exactly those terms (device-resident, deterministic reductions, capturable in the tCG graph).
// The ADMM step (:71-88: y, As, Af, x <- x - sigma*As, w <- w - sigma*Af, eX = x + bA, z, obj) is msdp_dual_kkt; the dual
// slack operator of the eigen step is then X = eX - diag(z), served by the dense-S path of affine.cu (h->eS, h->zdiag).
#include <math.h>
#include <string.h>
#include <algorithm>
#include "affine.h"
#include "gemm.h"
#include "kernels.cuh"
#include "rowops.cuh"

#define DG_T 32  // Gram tile
#define DG_R 64  // rows staged per step

// ---- kernels ------------------------------------------------------------------------------------------------------------
// part[chunk][i][j] = sum over the rows of the chunk of A[r, i] * B[r, j]   (i, j < ld; 32 x 32 tile per block)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_gram(const double* __restrict__ A, const double* __restrict__ B, int ld, int64_t nrows, int nchunks,
                double* __restrict__ part, const int* pred, RtrState* st, int skip_if_stopped) {
  __shared__ double sa[DG_R][DG_T + 1], sb[DG_R][DG_T + 1];
  if (pred && *pred == 0) return;
  if (skip_if_stopped && st->stop != 0) return;
  const int ti = blockIdx.x, tj = blockIdx.y, chunk = blockIdx.z, tid = threadIdx.x;
  const int64_t rows_per = (nrows + nchunks - 1) / nchunks;
  const int64_t rbeg = rows_per * chunk, rend = min(nrows, rbeg + rows_per);
  double acc[4] = {0, 0, 0, 0};
  const int oi = tid / 32, oj = tid % 32;
  for (int64_t r0 = rbeg; r0 < rend; r0 += DG_R) {
    const int tr = (int)min((int64_t)DG_R, rend - r0);
    for (int i = tid; i < DG_R * DG_T; i += blockDim.x) {
      const int r = i / DG_T, c = i % DG_T;
      const int ca = ti * DG_T + c, cb = tj * DG_T + c;
      sa[r][c] = (r < tr && ca < ld) ? A[(size_t)(r0 + r) * ld + ca] : 0.0;
      sb[r][c] = (r < tr && cb < ld) ? B[(size_t)(r0 + r) * ld + cb] : 0.0;
    }
    __syncthreads();
    for (int r = 0; r < DG_R; ++r) {
      const double bv = sb[r][oj];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fma(sa[r][oi + 8 * k], bv, acc[k]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = ti * DG_T + oi + 8 * k, j = tj * DG_T + oj;
    if (i < ld && j < ld) part[((size_t)chunk * ld + i) * ld + j] = acc[k];
  }
}
// out[i][j] = sum_chunks part[.][i][j] (+ the transposed entry when sym: out = G + G')
__global__ void k_dual_gram_sum(const double* __restrict__ part, double* __restrict__ out, int ld, int nchunks, int sym,
                                const int* pred, RtrState* st, int skip_if_stopped) {
  if (pred && *pred == 0) return;
  if (skip_if_stopped && st->stop != 0) return;
  const int nent = ld * ld;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nent; e += gridDim.x * blockDim.x) {
    const int i = e / ld, j = e - i * ld;
    double t = 0.0;
    for (int c = 0; c < nchunks; ++c) t += part[(size_t)c * nent + e];
    if (sym) {
      double u = 0.0;
      for (int c = 0; c < nchunks; ++c) u += part[(size_t)c * nent + (size_t)j * ld + i];
      t += u;
    }
    out[e] = t;
  }
}

// out[r, c] += 2 sigma (sum_q A1[r, q] M1[q, c] + sum_q A2[r, q] M2[q, c])   (A2 may be null)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_apply(double* __restrict__ out, const double* __restrict__ A1, const double* __restrict__ M1,
                 const double* __restrict__ A2, const double* __restrict__ M2, int ld, int64_t nrows, RtrState* st,
                 const int* pred, int skip_if_stopped) {
  if (pred && *pred == 0) return;
  if (skip_if_stopped && st->stop != 0) return;
  const double s2 = 2.0 * st->dual_sigma;
  const int64_t total = nrows * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / ld;
    const int c = (int)(i - r * ld);
    double acc = 0.0;
    const double* a1 = A1 + (size_t)r * ld;
    for (int q = 0; q < ld; ++q) acc = fma(a1[q], __ldg(M1 + (size_t)q * ld + c), acc);
    if (A2) {
      const double* a2 = A2 + (size_t)r * ld;
      for (int q = 0; q < ld; ++q) acc = fma(a2[q], __ldg(M2 + (size_t)q * ld + c), acc);
    }
    out[i] = fma(s2, acc, out[i]);
  }
}

// one block: st->tmp[6] = sigma/2 |G|_F^2 + sigma/2 sum_f Af_f^2 + k0,   Af = B'y - c_f - w/sigma,  y_k = isd_k r_k
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_cost_extra(const double* __restrict__ G, int ld, const double* __restrict__ r, const double* __restrict__ isd,
                      const int* __restrict__ Bjc, const int* __restrict__ Bir, const double* __restrict__ Bpr,
                      const double* __restrict__ cf, const double* __restrict__ wf, int nfree, RtrState* st) {
  __shared__ double sm[32];
  const double sigma = st->dual_sigma;
  double q[1] = {0.0};
  for (int e = threadIdx.x; e < ld * ld; e += blockDim.x) q[0] = fma(G[e], G[e], q[0]);
  block_sum<1>(q, sm);
  double total = 0.0;
  if (threadIdx.x == 0) total = q[0];
  double af2 = 0.0;
  for (int f = 0; f < nfree; ++f) {
    double v[1] = {0.0};
    for (int e = Bjc[f] + threadIdx.x; e < Bjc[f + 1]; e += blockDim.x) {
      const int k = Bir[e];
      v[0] = fma(Bpr[e], isd[k] * r[k], v[0]);
    }
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) {
      const double af = v[0] - cf[f] - wf[f] / sigma;
      af2 = fma(af, af, af2);
    }
  }
  if (threadIdx.x == 0) st->tmp[6] = 0.5 * sigma * total + 0.5 * sigma * af2 + st->dual_k0;
}

// C_eff = bA + x - sigma*c -> Cd and eS ; partial sums of <bA + x, c>, |c|^2, |x|^2 -> st->tmp[0..2]
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_refresh(const double* __restrict__ bA, const double* __restrict__ x, const double* __restrict__ c, double sigma,
                   double* __restrict__ Cd, double* __restrict__ eS, int64_t nn, RtrState* st, double* partials) {
  __shared__ double sm[3 * 32];
  double q[3] = {0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const double u = bA[i] + x[i], cv = c[i];
    const double v = u - sigma * cv;
    Cd[i] = v;
    eS[i] = v;
    q[0] = fma(u, cv, q[0]);
    q[1] = fma(cv, cv, q[1]);
    q[2] = fma(x[i], x[i], q[2]);
  }
  double tot[3];
  if (grid_sum_last<3>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      st->dual_k0 = -tot[0] + 0.5 * sigma * tot[1] + tot[2] / (2.0 * sigma);
      st->dual_sigma = sigma;
    }
  }
}

// ADMM step on the n x n arrays (ManiDSDP_unitdiag.m:74,78,80,86): As = T - S + c ; x <- x - sigma*As ; eX = x + bA
// tmp[0] = |As|^2, tmp[1] = <c, eX>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_admm(const double* __restrict__ T, const double* __restrict__ S, const double* __restrict__ c,
                const double* __restrict__ bA, double* __restrict__ x, double* __restrict__ eX, double sigma, int update,
                int64_t nn, RtrState* st, double* partials) {
  __shared__ double sm[2 * 32];
  double q[2] = {0, 0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const double as = T[i] - S[i] + c[i];
    const double xn = x[i] - sigma * as;
    if (update) x[i] = xn;
    const double ex = xn + bA[i];
    eX[i] = ex;
    q[0] = fma(as, as, q[0]);
    q[1] = fma(c[i], ex, q[1]);
  }
  double tot[2];
  if (grid_sum_last<2>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      st->tmp[0] = tot[0];
      st->tmp[1] = tot[1];
    }
  }
}

// y_k = isd_k r_k (:73) ; tmp[2] = b'y (:77)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_y(const double* __restrict__ r, const double* __restrict__ isd, const double* __restrict__ b,
             double* __restrict__ y, int64_t m, RtrState* st, double* partials) {
  __shared__ double sm[32];
  double q[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double v = isd[k] * r[k];
    y[k] = v;
    q[0] = fma(b[k], v, q[0]);
  }
  double tot[1];
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) st->tmp[2] = tot[0];
  }
}

// one block: Af = B'y - c_f (:75), w <- w - sigma*Af (:79) ; tmp[3] = |Af|^2, tmp[5] = c_f'w
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_free(const double* __restrict__ y, const int* __restrict__ Bjc, const int* __restrict__ Bir,
                const double* __restrict__ Bpr, const double* __restrict__ cf, double* __restrict__ wf, int nfree,
                double sigma, int update, RtrState* st) {
  __shared__ double sm[32];
  double af2 = 0.0, cw = 0.0;
  for (int f = 0; f < nfree; ++f) {
    double v[1] = {0.0};
    for (int e = Bjc[f] + threadIdx.x; e < Bjc[f + 1]; e += blockDim.x) v[0] = fma(Bpr[e], y[Bir[e]], v[0]);
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) {
      const double af = v[0] - cf[f];
      const double wn = wf[f] - sigma * af;
      if (update) wf[f] = wn;
      af2 = fma(af, af, af2);
      cw = fma(cf[f], wn, cw);
    }
  }
  if (threadIdx.x == 0) {
    st->tmp[3] = af2;
    st->tmp[5] = cw;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
static int dgrid(const manisdp_handle* h, int64_t total) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->num_sms * 8, (total + MSDP_THREADS - 1) / MSDP_THREADS));
}

template <typename T>
static int dual_to_dev(manisdp_handle* h, T** dst, const T* src, size_t count) {
  CUDA_TRY(h, cudaMalloc((void**)dst, std::max<size_t>(1, count) * sizeof(T)));
  if (count) CUDA_TRY(h, cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
  return MANISDP_OK;
}

// Gram buffers follow the row length of the work arrays
static int dual_gram_buffers(manisdp_handle* h) {
  DualData& d = h->dual;
  const int ld = (int)h->ld;
  if (d.gram_ld >= ld && d.gram[0]) return MANISDP_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const int cap = std::min<int>(MSDP_MAX_LD_AFFINE, ld + ld / 2 + 16);
  double** bufs[] = {&d.gram[0], &d.gram[1], &d.gramS, &d.gramW};
  for (double** b : bufs) {
    if (*b) cudaFree(*b);
    *b = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)b, (size_t)cap * cap * sizeof(double)));
  }
  d.gram_ld = cap;
  return MANISDP_OK;
}

static int dual_nchunks(const manisdp_handle* h) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(32, h->n / DG_R));
}

// out (ld x ld) = A' B (+ transposed when sym)
static int dual_gram(manisdp_handle* h, const double* A, const double* B, double* out, int sym, const int* pred, int skip) {
  DualData& d = h->dual;
  const int ld = (int)h->ld, nch = dual_nchunks(h);
  const size_t need = (size_t)nch * ld * ld * sizeof(double);
  if (need > d.gram_part_cap) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (d.gram_part) cudaFree(d.gram_part);
    d.gram_part = nullptr;
    d.gram_part_cap = 0;
    CUDA_TRY(h, cudaMalloc((void**)&d.gram_part, 2 * need));
    d.gram_part_cap = 2 * need;
  }
  const int nt = (ld + DG_T - 1) / DG_T;
  dim3 grid(nt, nt, nch);
  k_dual_gram<<<grid, MSDP_THREADS, 0, h->stream>>>(A, B, ld, h->n, nch, d.gram_part, pred, h->st, skip);
  KERNEL_CHECK(h);
  k_dual_gram_sum<<<std::max(1, (ld * ld + 255) / 256), 256, 0, h->stream>>>(d.gram_part, out, ld, nch, sym, pred, h->st,
                                                                          skip);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_dual_setup(manisdp_handle* h, const manisdp_problem* pb, const std::vector<double>& dAAt) {
  DualData& d = h->dual;
  const int64_t n = h->n, m = h->m, nn = n * n;
  if (h->s_mode != MODE_DENSE) return msdp_fail(h, MANISDP_E_ARG, "dual handles need the dense S representation");
  d.on = 1;
  h->rank_strict = 1;
  std::vector<double> isd((size_t)m);
  for (int64_t k = 0; k < m; ++k) {
    if (!(dAAt[(size_t)k] > 0.0)) return msdp_fail(h, MANISDP_E_ARG, "dual: diag(A*A') must be positive (empty constraint row)");
    isd[(size_t)k] = 1.0 / sqrt(dAAt[(size_t)k]);
  }
  MSDP_TRY(dual_to_dev(h, &d.isd, isd.data(), (size_t)m));
  MSDP_TRY(dual_to_dev(h, &d.borig, pb->b, (size_t)m));
  // bA = A' D^-1 b (ManiDSDP_unitdiag.m:43-44), dense n x n in the column-major vec order of the (symmetric) input
  std::vector<double> bA((size_t)nn, 0.0);
  for (int64_t k = 0; k < m; ++k) {
    const double s = pb->b[k] / dAAt[(size_t)k];
    if (s == 0.0) continue;
    for (uint64_t e = pb->At_jc[k]; e < pb->At_jc[k + 1]; ++e) bA[(size_t)pb->At_ir[e]] += pb->At_pr[e] * s;
  }
  MSDP_TRY(dual_to_dev(h, &d.bA, bA.data(), (size_t)nn));
  CUDA_TRY(h, cudaMalloc((void**)&d.cpsd, (size_t)nn * sizeof(double)));
  CUDA_TRY(h, cudaMemcpy(d.cpsd, h->Cdense, (size_t)nn * sizeof(double), cudaMemcpyDeviceToDevice));
  CUDA_TRY(h, cudaMalloc((void**)&d.x, (size_t)nn * sizeof(double)));
  CUDA_TRY(h, cudaMemset(d.x, 0, (size_t)nn * sizeof(double)));
  if (!h->Mbuf) CUDA_TRY(h, cudaMalloc((void**)&h->Mbuf, (size_t)nn * sizeof(double)));
  if (!h->Tbuf) {
    CUDA_TRY(h, cudaMalloc((void**)&h->Tbuf, (size_t)nn * sizeof(double)));
    CUDA_TRY(h, cudaMemset(h->Tbuf, 0, (size_t)nn * sizeof(double)));
  }
  // free part
  d.nfree = pb->nfree;
  double cnorm2 = 0.0;
  std::vector<double> ch((size_t)nn);
  CUDA_TRY(h, cudaMemcpy(ch.data(), d.cpsd, (size_t)nn * sizeof(double), cudaMemcpyDeviceToHost));
  for (double v : ch) cnorm2 += v * v;
  std::vector<int> bjc((size_t)d.nfree + 1, 0), bir;
  std::vector<double> bpr, cf((size_t)d.nfree, 0.0), w0((size_t)d.nfree, 0.0);
  if (d.nfree > 0) {
    if (!pb->B_jc || !pb->B_ir || !pb->B_pr || !pb->cf) return msdp_fail(h, MANISDP_E_ARG, "dual: nfree > 0 needs B (CSC) and cf");
    for (int64_t f = 0; f <= d.nfree; ++f) bjc[(size_t)f] = (int)pb->B_jc[f];
    const uint64_t nz = pb->B_jc[d.nfree];
    bir.resize((size_t)nz);
    bpr.assign(pb->B_pr, pb->B_pr + nz);
    for (uint64_t e = 0; e < nz; ++e) {
      if (pb->B_ir[e] >= (uint64_t)m) return msdp_fail(h, MANISDP_E_ARG, "dual: B row index out of range");
      bir[(size_t)e] = (int)pb->B_ir[e];
    }
    for (int64_t f = 0; f < d.nfree; ++f) {
      cf[(size_t)f] = pb->cf[f];
      cnorm2 += pb->cf[f] * pb->cf[f];
    }
  }
  d.normc = 1.0 + sqrt(cnorm2);  // :33 (c includes the free part)
  MSDP_TRY(dual_to_dev(h, &d.B_jc, bjc.data(), bjc.size()));
  MSDP_TRY(dual_to_dev(h, &d.B_ir, bir.data(), bir.size()));
  MSDP_TRY(dual_to_dev(h, &d.B_pr, bpr.data(), bpr.size()));
  MSDP_TRY(dual_to_dev(h, &d.cf, cf.data(), cf.size()));
  MSDP_TRY(dual_to_dev(h, &d.w, w0.data(), w0.size()));
  d.dirty = 1;
  return MANISDP_OK;
}

void msdp_dual_free(manisdp_handle* h) {
  DualData& d = h->dual;
  void* ptrs[] = {d.bA, d.cpsd, d.x, d.isd, d.borig, d.gram[0], d.gram[1], d.gramS, d.gramW, d.gram_part,
                  d.B_jc, d.B_ir, d.B_pr, d.cf, d.w};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  d = DualData();
}

int msdp_dual_refresh(manisdp_handle* h) {
  DualData& d = h->dual;
  const int64_t nn = h->n * h->n;
  k_dual_refresh<<<dgrid(h, nn), MSDP_THREADS, 0, h->stream>>>(d.bA, d.x, d.cpsd, h->sigma, h->Cdense, h->eS, nn, h->st,
                                                             h->partials);
  KERNEL_CHECK(h);
  d.dirty = 0;
  h->cache_valid = h->grad_valid = 0;  // (captured graphs stay valid: k0 / sigma are read from the device state, a new
                                       //  sigma rebuilds them through graph_sigma in rtr.cu)
  return MANISDP_OK;
}

// w >= 0: the Gram matrix of the point buffer w is kept for the gradient / Hessian at that point; w < 0: scratch
int msdp_dual_cost_extra(manisdp_handle* h, const double* Z, const double* resid, int w) {
  DualData& d = h->dual;
  MSDP_TRY(dual_gram_buffers(h));
  double* G = w >= 0 ? d.gram[w] : d.gramS;
  MSDP_TRY(dual_gram(h, Z, Z, G, 0, nullptr, 0));
  k_dual_cost_extra<<<1, MSDP_THREADS, 0, h->stream>>>(G, (int)h->ld, resid, d.isd, d.B_jc, d.B_ir, d.B_pr, d.cf, d.w,
                                                      (int)d.nfree, h->st);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_dual_grad_extra(manisdp_handle* h, const double* Z, double* G, int w, const int* pred) {
  DualData& d = h->dual;
  k_dual_apply<<<dgrid(h, h->n * h->ld), MSDP_THREADS, 0, h->stream>>>(G, Z, d.gram[w], nullptr, nullptr, (int)h->ld, h->n,
                                                                      h->st, pred, 0);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_dual_hess_extra(manisdp_handle* h, const double* Y, const double* D, double* Hout, int w, int skip) {
  DualData& d = h->dual;
  MSDP_TRY(dual_gram(h, Y, D, d.gramW, 1, nullptr, skip));  // W = Y'U + U'Y
  k_dual_apply<<<dgrid(h, h->n * h->ld), MSDP_THREADS, 0, h->stream>>>(Hout, Y, d.gramW, D, d.gram[w], (int)h->ld, h->n,
                                                                      h->st, nullptr, skip);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// ADMM step + KKT residues (ManiDSDP_unitdiag.m:71-89); leaves eX in h->eS and z in h->zdiag for the eigen step
int msdp_dual_kkt(manisdp_handle* h, int update, manisdp_kkt_info* out) {
  DualData& d = h->dual;
  h->cache_valid = 0;
  MSDP_TRY(msdp_ensure_costgrad(h));  // r = A~ (vec(S) - c) at the current point
  const int w = h->pt;
  const double* Y = h->Ybuf[w];
  const int ld = (int)h->ld;
  const int64_t n = h->n, nn = n * n;
  k_dual_y<<<dgrid(h, h->m), MSDP_THREADS, 0, h->stream>>>(h->resid[w], d.isd, d.borig, h->y, h->m, h->st, h->partials);
  KERNEL_CHECK(h);
  k_dual_free<<<1, MSDP_THREADS, 0, h->stream>>>(h->y, d.B_jc, d.B_ir, d.B_pr, d.cf, d.w, (int)d.nfree, h->sigma, update,
                                                h->st);
  KERNEL_CHECK(h);
  MSDP_TRY(msdp_gemm_nt(h, Y, ld, Y, ld, (int)n, ld, h->Mbuf, 1.0, nullptr));  // S = Y Y'
  MSDP_TRY(msdp_affine_touch(h, h->resid[w], 1.0, nullptr, h->Tbuf));          // T = A~' r = P sc
  k_dual_admm<<<dgrid(h, nn), MSDP_THREADS, 0, h->stream>>>(h->Tbuf, h->Mbuf, d.cpsd, d.bA, d.x, h->eS, h->sigma, update,
                                                          nn, h->st, h->partials);
  KERNEL_CHECK(h);
  // z = sum(S.*eX) = rowwise <Y_a, (eX Y)_a>  (:81), X = eX - diag(z) is applied as (eS, zdiag) by msdp_affine_apply_S
  MSDP_TRY(msdp_gemm_nn(h, h->eS, (int)n, Y, ld, ld, h->Hd, ld, 1.0, 0.0, nullptr));
  MSDP_TRY(msdp_affine_rowdot(h, Y, h->Hd, h->zdiag, 4));
  const int keep = h->pt;
  MSDP_TRY(msdp_sync_state(h));
  h->pt = keep;
  const RtrState* s = h->st_host;
  const double zsum = s->tmp[4];
  out->pinf = (sqrt(s->tmp[0]) + sqrt(s->tmp[3])) / d.normc;  // :76
  out->by = s->tmp[2];                                         // :77
  out->obj = s->tmp[1] + s->tmp[5] + zsum;                     // :86
  out->z_sum = zsum;
  out->gap = fabs(out->obj - out->by) / (1.0 + fabs(out->obj) + fabs(out->by));  // :88
  h->zshift = 0.0;
  h->y_kkt = h->y;
  d.dirty = 1;  // x (and eS) changed: C_eff / k0 are rebuilt before the next closure call
  h->cache_valid = h->grad_valid = 0;
  return MANISDP_OK;
}

// ---- state access (tests, warm starts) -------------------------------------------------------------------------------
extern "C" int manisdp_dual_get_state(manisdp_t* h, double* x, double* w) {
  if (!h || !h->dual.on) return msdp_fail(h, MANISDP_E_ARG, "dual_get_state: dual handle needed");
  CUDA_TRY(h, cudaSetDevice(h->device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (x) CUDA_TRY(h, cudaMemcpy(x, h->dual.x, (size_t)h->n * h->n * sizeof(double), cudaMemcpyDeviceToHost));
  if (w && h->dual.nfree) CUDA_TRY(h, cudaMemcpy(w, h->dual.w, (size_t)h->dual.nfree * sizeof(double), cudaMemcpyDeviceToHost));
  return MANISDP_OK;
}

extern "C" int manisdp_dual_set_state(manisdp_t* h, const double* x, const double* w) {
  if (!h || !h->dual.on) return msdp_fail(h, MANISDP_E_ARG, "dual_set_state: dual handle needed");
  CUDA_TRY(h, cudaSetDevice(h->device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (x) CUDA_TRY(h, cudaMemcpy(h->dual.x, x, (size_t)h->n * h->n * sizeof(double), cudaMemcpyHostToDevice));
  if (w && h->dual.nfree) CUDA_TRY(h, cudaMemcpy(h->dual.w, w, (size_t)h->dual.nfree * sizeof(double), cudaMemcpyHostToDevice));
  h->dual.dirty = 1;
  h->cache_valid = h->grad_valid = 0;
  return MANISDP_OK;
}
