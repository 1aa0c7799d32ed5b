// affine.cu -- closures of the three affine drivers on  f(Y) = <C, YY'> + sigma/2 |A(YY') - b - y/sigma|^2 :
//   ManiSDP_unitdiag.m:152-171 (oblique), ManiSDP_unittrace.m:156-177 (sphere), ManiSDP.m:149-165 (Euclidean),
// plus their KKT step (ManiSDP_unitdiag.m:59-71) and the dual-slack operator for the eigen step.
//
// The reference forms X = Y'Y, YU = Y'U and AyU = mat(At*(A*vec(YU))) densely for every problem.  Here the constraint
// operator has two device representations chosen at create time from the sparsity of At:
//   A sparse (theta-like, nnz(At) << n^2):  K2 `sddmm`   w_k = sum_{(i,j) in A_k} a <P_i, Q_j>        (never forms n x n)
//                                           K3 `rowlist` H_i += sum_{(j,k) in row i} a (c1 v1_k V1_j + c2 v2_k V2_j)
//   A dense  (BQP / quartic sphere, pattern covers X):  M = P Q' by FP64 DMMA GEMM, w = gather over the CSC pattern,
//                                           T = scatter to the touched positions, H += T*Y by GEMM   (K4)
// and so has the matrix S = C + sigma*At(r):  dense n x n (GEMM) when C or the pattern is dense, CSR + rowlist otherwise.
// Index arithmetic: At rows are 64-bit linear indices r = j*n + i, split with integer div/mod at create time (bit-exact).
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include "affine.h"
#include "gemm.h"
#include "kernels.cuh"
#include "rowops.cuh"
#include "scalar_logic.cuh"

// scratch scalar slots in st->tmp used by this file
#define T_F 0      // cost
#define T_G2 1     // |grad|^2
#define T_RR 2     // |r|^2
#define T_CX 3     // <C, X>
#define T_S 4      // sphere scalars
#define T_AUX 5

// penalty the closure kernels see: sigma of the primal AL drivers; -sigma on a dual (ManiDSDP) handle, whose cost is
// <C_eff, S> + sigma/2 |S|^2 - sigma/2 |A~ vec(S) - A~ c|^2 + const (dual.cu)
static inline double csig(const manisdp_handle* h) { return h->dual.on ? -h->sigma : h->sigma; }

static int mgrid(const manisdp_handle* h, int64_t total, int per_block = MSDP_THREADS) {
  int64_t nb = (total + per_block - 1) / per_block;
  const int64_t cap = std::min<int64_t>((int64_t)h->num_sms * 8, MSDP_MAX_BLOCKS);
  return (int)std::max<int64_t>(1, std::min(nb, cap));
}

// ======================================================================================================================
// kernels
// ======================================================================================================================

// K2 (sparse A): w_k = sum_e a_e <P_{i_e}, Q_{j_e}> ; mode 1 additionally r_k = w_k - b_k - y_k/sigma and sum r_k^2
// The entries of a constraint are taken four at a time: indices first, then the eight operand rows, then the products --
// the gathers of a batch are independent and in flight together (summation order unchanged).
template <int GS, int VPL>
__device__ __forceinline__ double sddmm_range(const int* __restrict__ ei, const int* __restrict__ ej,
                                              const double* __restrict__ ea, const double* __restrict__ P,
                                              const double* __restrict__ Q, int ld, int e0, int e1, int gl) {
  const int nvec = ld / 2;
  double acc = 0.0;
  int e = e0;
  for (; e + 4 <= e1; e += 4) {
    double d[4] = {0.0, 0.0, 0.0, 0.0};
    const double* pi[4];
    const double* qj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pi[u] = P + (size_t)ei[e + u] * ld;
      qj[u] = Q + (size_t)ej[e + u] * ld;
    }
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        double2 a[4], bq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = ldg2(pi[u] + 2 * c);
          bq[u] = ldg2(qj[u] + 2 * c);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) d[u] += a[u].x * bq[u].x + a[u].y * bq[u].y;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = fma(ea[e + u], d[u], acc);
  }
  for (; e < e1; ++e) {
    const double* p1 = P + (size_t)ei[e] * ld;
    const double* q1 = Q + (size_t)ej[e] * ld;
    double d = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        const double2 a = ldg2(p1 + 2 * c), bq = ldg2(q1 + 2 * c);
        d += a.x * bq.x + a.y * bq.y;
      }
    }
    acc = fma(ea[e], d, acc);
  }
  return acc;
}

template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_sddmm(const int* __restrict__ kptr, const int* __restrict__ ei, const int* __restrict__ ej,
            const double* __restrict__ ea, const double* __restrict__ P, const double* __restrict__ Q, int ld, int64_t m,
            double* __restrict__ out, int mode, const double* __restrict__ b, const double* __restrict__ y,
            double inv_sigma, RtrState* st, double* partials, int skip_if_stopped) {
  __shared__ double sm[32];
  if (skip_if_stopped && st->stop != 0) return;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  double q[1] = {0.0};
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; k < m; k += ngroups) {
    double acc = sddmm_range<GS, VPL>(ei, ej, ea, P, Q, ld, kptr[k], kptr[k + 1], gl);
    acc = group_sum<GS>(acc, mask);
    if (mode == 1) acc = acc - b[k] - y[k] * inv_sigma;
    if (gl == 0) {
      out[k] = acc;
      if (mode == 1) q[0] += acc * acc;
    }
  }
  if (mode == 1) {
    double tot[1];
    __syncwarp();
    if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
      if (threadIdx.x == 0) st->tmp[T_RR] = tot[0];
    }
  }
}

// segmented form (some constraint is long): one row group per SEGMENT of at most SDDMM_SEG entries ...
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_sddmm_seg(const int* __restrict__ sptr, const int* __restrict__ ei, const int* __restrict__ ej,
                const double* __restrict__ ea, const double* __restrict__ P, const double* __restrict__ Q, int ld,
                int64_t nseg, double* __restrict__ segval, RtrState* st, int skip_if_stopped) {
  if (skip_if_stopped && st->stop != 0) return;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t sg = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; sg < nseg; sg += ngroups) {
    double acc = sddmm_range<GS, VPL>(ei, ej, ea, P, Q, ld, sptr[sg], sptr[sg + 1], gl);
    acc = group_sum<GS>(acc, mask);
    if (gl == 0) segval[sg] = acc;
  }
}
// ... and the per-constraint sums of the segment values in segment order (deterministic), with the mode-1 residual
__global__ void __launch_bounds__(MSDP_THREADS)
    k_sddmm_finish(const int* __restrict__ ksegs, const double* __restrict__ segval, int64_t m, double* __restrict__ out,
                   int mode, const double* __restrict__ b, const double* __restrict__ y, double inv_sigma, RtrState* st,
                   double* partials, int skip_if_stopped) {
  __shared__ double sm[32];
  if (skip_if_stopped && st->stop != 0) return;
  double q[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    double acc = 0.0;
    for (int sg = ksegs[k]; sg < ksegs[k + 1]; ++sg) acc += segval[sg];
    if (mode == 1) {
      acc = acc - b[k] - y[k] * inv_sigma;
      q[0] += acc * acc;
    }
    out[k] = acc;
  }
  if (mode == 1) {
    double tot[1];
    if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
      if (threadIdx.x == 0) st->tmp[T_RR] = tot[0];
    }
  }
}

// dense A: w_k = sum_e a_e M[lin_e]  (A * vec(M), ManiSDP_unitdiag.m:154,168), same modes as k_sddmm
__global__ void __launch_bounds__(MSDP_THREADS)
    k_gather_spmv(const int* __restrict__ kptr, const int* __restrict__ klin, const double* __restrict__ ka,
                  const double* __restrict__ M, int64_t m, double* __restrict__ out, int mode,
                  const double* __restrict__ b, const double* __restrict__ y, double inv_sigma, RtrState* st,
                  double* partials, int skip_if_stopped) {
  __shared__ double sm[32];
  if (skip_if_stopped && st->stop != 0) return;
  double q[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;  // a multiple of 32: the lanes of a warp walk 32 neighbouring constraints
  const int lane = threadIdx.x & 31;
  for (int64_t kb = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; kb < m; kb += stride) {
    const int64_t k = kb + lane;
    int e0 = 0, e1 = 0;
    if (k < m) {
      e0 = kptr[k];
      e1 = kptr[k + 1];
    }
    // short constraints: one lane each.  LONG ones (a SOS constraint collects every position of S that represents its
    // monomial: 1831 entries for the constant term of BQP-60, ~60 for a quadratic one) are walked by the whole warp, lane
    // partial sums joined by the fixed butterfly -- the kernel used to take as long as its longest constraint (335 us).
    const bool is_long = (e1 - e0) > 32;
    double acc = 0.0;
    if (!is_long)
      for (int e = e0; e < e1; ++e) acc = fma(__ldg(ka + e), __ldg(M + klin[e]), acc);
    unsigned longmask = __ballot_sync(0xffffffffu, is_long);
    while (longmask) {
      const int src = __ffs(longmask) - 1;
      longmask &= longmask - 1;
      const int s0 = __shfl_sync(0xffffffffu, e0, src), s1 = __shfl_sync(0xffffffffu, e1, src);
      double part = 0.0;
      for (int e = s0 + lane; e < s1; e += 32) part = fma(__ldg(ka + e), __ldg(M + klin[e]), part);
      part = warp_sum(part);
      if (lane == src) acc = part;
    }
    if (k < m) {
      if (mode == 1) {
        acc = acc - b[k] - y[k] * inv_sigma;
        q[0] += acc * acc;
      }
      out[k] = acc;
    }
  }
  if (mode == 1) {
    double tot[1];
    if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
      if (threadIdx.x == 0) st->tmp[T_RR] = tot[0];
    }
  }
}

// st->tmp[slot] = sum_i a[i]*b[i], i < len
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dot_len(const double* __restrict__ a, const double* __restrict__ b, int64_t len, RtrState* st, double* partials,
              int slot) {
  __shared__ double sm[32];
  double q[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) q[0] = fma(a[i], b[i], q[0]);
  double tot[1];
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) st->tmp[slot] = tot[0];
  }
}

// touched positions of the n x n matrix: dst[pos] = (base ? base[pos] : 0) + coef * sum_e la_e * vec[lk_e]
__global__ void __launch_bounds__(MSDP_THREADS)
    k_touch_update(const int* __restrict__ upos, const int* __restrict__ lptr, const int* __restrict__ lk,
                   const double* __restrict__ la, const double* __restrict__ vec, double coef,
                   const double* __restrict__ base, double* __restrict__ dst, int64_t nu, const int* pred,
                   RtrState* st, int skip_if_stopped) {
  if (pred && *pred == 0) return;
  if (skip_if_stopped && st->stop != 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += stride) {
    const int e0 = lptr[u], e1 = lptr[u + 1];
    double acc = 0.0;
    for (int e = e0; e < e1; ++e) acc = fma(__ldg(la + e), __ldg(vec + lk[e]), acc);
    const int pos = upos[u];
    dst[pos] = (base ? base[pos] : 0.0) + coef * acc;
  }
}

// K3 (sparse): out_i = beta*out_i + alphaC * sum_{j in C_i} C_ij V1_j
//                      + sum_{(j,k,a) in rowlist i} a * (c1*vec1_k * V1_j + c2*vec2_k * V2_j)
struct RowlistArgs {
  const int *crowptr, *ccol;
  const double* cval;  // may be null (dense S: no CSR part)
  double alphaC;
  const int *rptr, *rj, *rk;
  const double* ra;
  const double *vec1, *vec2;
  double c1, c2;
  const double *V1, *V2;
  double* out;
  double beta;
  int64_t nrows;
  int ld;
  const int* pred;
  int skip_if_stopped;
};
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS) k_rowlist_apply(RowlistArgs a, RtrState* st) {
  if (a.pred && *a.pred == 0) return;
  if (a.skip_if_stopped && st->stop != 0) return;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = a.ld / 2, ld = a.ld;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < a.nrows; row += ngroups) {
    double2 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) acc[t] = make_double2(0.0, 0.0);
    if (a.cval) {
      for (int e = a.crowptr[row]; e < a.crowptr[row + 1]; ++e) {
        const double w = a.alphaC * __ldg(a.cval + e);
        const double* pj = a.V1 + (size_t)__ldg(a.ccol + e) * ld;
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int c = gl + GS * t;
          if (c < nvec) {
            const double2 u = ldg2(pj + 2 * c);
            acc[t].x = fma(w, u.x, acc[t].x);
            acc[t].y = fma(w, u.y, acc[t].y);
          }
        }
      }
    }
    if (a.ra) {
      // The entries of a row are taken GS at a time: lane l of the row group loads entry eb + l (coalesced index / value
      // loads, then the multiplier gathers vec[k]) and the group walks the batch through shuffles, four entries per step
      // with all their operand-row gathers in flight together.  Summation order is the entry order (V1 term, then V2
      // term), exactly as a one-by-one walk -- results do not depend on the batching.  Long rows (a dense block of a
      // multi-block moment relaxation has hundreds of entries per row) used to cost three dependent round trips per entry.
      const int e0 = a.rptr[row], e1 = a.rptr[row + 1];
      for (int eb = e0; eb < e1; eb += GS) {
        const int e = eb + gl;
        int j = 0;
        double w1 = 0.0, w2 = 0.0;
        if (e < e1) {
          const double av = __ldg(a.ra + e);
          j = __ldg(a.rj + e);
          const int k = __ldg(a.rk + e);
          if (a.vec1) w1 = a.c1 * av * __ldg(a.vec1 + k);
          if (a.vec2) w2 = a.c2 * av * __ldg(a.vec2 + k);
        }
        const int cnt = min(GS, e1 - eb);
        int u = 0;
        for (; u + 4 <= cnt; u += 4) {
          size_t jo[4];
          double a1[4], a2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            jo[q] = (size_t)__shfl_sync(mask, j, u + q, GS) * ld;
            a1[q] = __shfl_sync(mask, w1, u + q, GS);
            a2[q] = __shfl_sync(mask, w2, u + q, GS);
          }
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int c = gl + GS * t;
            if (c < nvec) {
              double2 x1[4], x2[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (a.vec1) x1[q] = ldg2(a.V1 + jo[q] + 2 * c);
                if (a.vec2) x2[q] = ldg2(a.V2 + jo[q] + 2 * c);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (a.vec1) {
                  acc[t].x = fma(a1[q], x1[q].x, acc[t].x);
                  acc[t].y = fma(a1[q], x1[q].y, acc[t].y);
                }
                if (a.vec2) {
                  acc[t].x = fma(a2[q], x2[q].x, acc[t].x);
                  acc[t].y = fma(a2[q], x2[q].y, acc[t].y);
                }
              }
            }
          }
        }
        for (; u < cnt; ++u) {
          const size_t jo = (size_t)__shfl_sync(mask, j, u, GS) * ld;
          const double b1 = __shfl_sync(mask, w1, u, GS), b2 = __shfl_sync(mask, w2, u, GS);
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int c = gl + GS * t;
            if (c < nvec) {
              if (a.vec1) {
                const double2 x = ldg2(a.V1 + jo + 2 * c);
                acc[t].x = fma(b1, x.x, acc[t].x);
                acc[t].y = fma(b1, x.y, acc[t].y);
              }
              if (a.vec2) {
                const double2 x = ldg2(a.V2 + jo + 2 * c);
                acc[t].x = fma(b2, x.x, acc[t].x);
                acc[t].y = fma(b2, x.y, acc[t].y);
              }
            }
          }
        }
      }
    }
    const size_t rb = (size_t)row * ld;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        if (a.beta != 0.0) {
          const double2 o = ld2(a.out + rb + 2 * c);
          acc[t].x = fma(a.beta, o.x, acc[t].x);
          acc[t].y = fma(a.beta, o.y, acc[t].y);
        }
        st2(a.out + rb + 2 * c, acc[t]);
      }
    }
  }
}

// K3 for LONG rows (dense blocks of a multi-block moment relaxation: hundreds of entries per row, few thousand rows): one
// CTA per row.  Warp w of the CTA takes the 32-entry batches w, w + 8, w + 16, ... of the row (same batched walk as
// above), the eight partial rows meet in shared memory and are added in warp order -- a fixed summation tree, so results
// are reproducible -- together with the C part, alpha / beta and the store.  One warp per row left the SMs a third full
// and every warp walking ~11 dependent batches; this form has 8x the loads in flight per row.
template <int VPL>
__global__ void __launch_bounds__(MSDP_THREADS) k_rowlist_apply_wide(RowlistArgs a, RtrState* st) {
  __shared__ double2 part[MSDP_THREADS / 32][MSDP_MAX_LD / 2];
  if (a.pred && *a.pred == 0) return;
  if (a.skip_if_stopped && st->stop != 0) return;
  constexpr int GS = 32, NW = MSDP_THREADS / 32;
  const int gl = threadIdx.x % 32, wid = threadIdx.x / 32, nvec = a.ld / 2, ld = a.ld;
  for (int64_t row = blockIdx.x; row < a.nrows; row += gridDim.x) {
    double2 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) acc[t] = make_double2(0.0, 0.0);
    if (a.cval && wid == 0) {
      for (int e = a.crowptr[row]; e < a.crowptr[row + 1]; ++e) {
        const double w = a.alphaC * __ldg(a.cval + e);
        const double* pj = a.V1 + (size_t)__ldg(a.ccol + e) * ld;
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int c = gl + GS * t;
          if (c < nvec) {
            const double2 u = ldg2(pj + 2 * c);
            acc[t].x = fma(w, u.x, acc[t].x);
            acc[t].y = fma(w, u.y, acc[t].y);
          }
        }
      }
    }
    const int e0 = a.rptr[row], e1 = a.rptr[row + 1];
    for (int eb = e0 + GS * wid; eb < e1; eb += GS * NW) {
      const int e = eb + gl;
      int j = 0;
      double w1 = 0.0, w2 = 0.0;
      if (e < e1) {
        const double av = __ldg(a.ra + e);
        j = __ldg(a.rj + e);
        const int k = __ldg(a.rk + e);
        if (a.vec1) w1 = a.c1 * av * __ldg(a.vec1 + k);
        if (a.vec2) w2 = a.c2 * av * __ldg(a.vec2 + k);
      }
      const int cnt = min(GS, e1 - eb);
      int u = 0;
      for (; u + 4 <= cnt; u += 4) {
        size_t jo[4];
        double a1[4], a2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          jo[q] = (size_t)__shfl_sync(0xffffffffu, j, u + q) * ld;
          a1[q] = __shfl_sync(0xffffffffu, w1, u + q);
          a2[q] = __shfl_sync(0xffffffffu, w2, u + q);
        }
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int c = gl + GS * t;
          if (c < nvec) {
            double2 x1[4], x2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (a.vec1) x1[q] = ldg2(a.V1 + jo[q] + 2 * c);
              if (a.vec2) x2[q] = ldg2(a.V2 + jo[q] + 2 * c);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (a.vec1) {
                acc[t].x = fma(a1[q], x1[q].x, acc[t].x);
                acc[t].y = fma(a1[q], x1[q].y, acc[t].y);
              }
              if (a.vec2) {
                acc[t].x = fma(a2[q], x2[q].x, acc[t].x);
                acc[t].y = fma(a2[q], x2[q].y, acc[t].y);
              }
            }
          }
        }
      }
      for (; u < cnt; ++u) {
        const size_t jo = (size_t)__shfl_sync(0xffffffffu, j, u) * ld;
        const double b1 = __shfl_sync(0xffffffffu, w1, u), b2 = __shfl_sync(0xffffffffu, w2, u);
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int c = gl + GS * t;
          if (c < nvec) {
            if (a.vec1) {
              const double2 x = ldg2(a.V1 + jo + 2 * c);
              acc[t].x = fma(b1, x.x, acc[t].x);
              acc[t].y = fma(b1, x.y, acc[t].y);
            }
            if (a.vec2) {
              const double2 x = ldg2(a.V2 + jo + 2 * c);
              acc[t].x = fma(b2, x.x, acc[t].x);
              acc[t].y = fma(b2, x.y, acc[t].y);
            }
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) part[wid][c] = acc[t];
    }
    __syncthreads();
    const size_t rb = (size_t)row * ld;
    for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
      double2 sres = part[0][c];
#pragma unroll
      for (int w = 1; w < NW; ++w) {
        sres.x += part[w][c].x;
        sres.y += part[w][c].y;
      }
      if (a.beta != 0.0) {
        const double2 o = ld2(a.out + rb + 2 * c);
        sres.x = fma(a.beta, o.x, sres.x);
        sres.y = fma(a.beta, o.y, sres.y);
      }
      st2(a.out + rb + 2 * c, sres);
    }
    __syncthreads();
  }
}

// The same for NARROW operands (ld < 64, i.e. fewer than 32 vectors per row: the first and the last outer iterations of
// a multi-block solve, where the factors are thin).  A warp splits into 32 / GS sub-groups of GS lanes; every sub-group
// takes its own entry of the 32-entry batch, so 32 / GS operand rows are gathered at once and four steps are in flight;
// the sub-group accumulators are joined by a butterfly over the lane bits above GS, the warps through shared memory.
template <int GS>
__global__ void __launch_bounds__(MSDP_THREADS) k_rowlist_apply_wide_narrow(RowlistArgs a, RtrState* st) {
  __shared__ double2 part[MSDP_THREADS / 32][32];
  if (a.pred && *a.pred == 0) return;
  if (a.skip_if_stopped && st->stop != 0) return;
  constexpr int NW = MSDP_THREADS / 32, NG = 32 / GS;
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32, gl = lane % GS, grp = lane / GS;
  const int nvec = a.ld / 2, ld = a.ld;
  const bool colok = gl < nvec;
  for (int64_t row = blockIdx.x; row < a.nrows; row += gridDim.x) {
    double2 acc = make_double2(0.0, 0.0);
    if (a.cval && wid == 0 && grp == 0) {
      for (int e = a.crowptr[row]; e < a.crowptr[row + 1]; ++e) {
        const double w = a.alphaC * __ldg(a.cval + e);
        if (colok) {
          const double2 u = ldg2(a.V1 + (size_t)__ldg(a.ccol + e) * ld + 2 * gl);
          acc.x = fma(w, u.x, acc.x);
          acc.y = fma(w, u.y, acc.y);
        }
      }
    }
    const int e0 = a.rptr[row], e1 = a.rptr[row + 1];
    for (int eb = e0 + 32 * wid; eb < e1; eb += 32 * NW) {
      const int e = eb + lane;
      int j = 0;
      double w1 = 0.0, w2 = 0.0;  // lanes past the end of the row carry zero weights and a valid row index (0)
      if (e < e1) {
        const double av = __ldg(a.ra + e);
        j = __ldg(a.rj + e);
        const int k = __ldg(a.rk + e);
        if (a.vec1) w1 = a.c1 * av * __ldg(a.vec1 + k);
        if (a.vec2) w2 = a.c2 * av * __ldg(a.vec2 + k);
      }
      const int cnt = min(32, e1 - eb);
      for (int u = 0; u < cnt; u += 4 * NG) {
        size_t jo[4];
        double a1[4], a2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int src = (u + q * NG + grp) & 31;  // entries past cnt have zero weights (w1 = w2 = 0 beyond the row)
          const bool live = (u + q * NG + grp) < 32;
          jo[q] = (size_t)__shfl_sync(0xffffffffu, j, src) * ld;
          a1[q] = __shfl_sync(0xffffffffu, w1, src);
          a2[q] = __shfl_sync(0xffffffffu, w2, src);
          if (!live) {
            a1[q] = 0.0;
            a2[q] = 0.0;
            jo[q] = 0;
          }
        }
        if (colok) {
          double2 x1[4], x2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (a.vec1) x1[q] = ldg2(a.V1 + jo[q] + 2 * gl);
            if (a.vec2) x2[q] = ldg2(a.V2 + jo[q] + 2 * gl);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (a.vec1) {
              acc.x = fma(a1[q], x1[q].x, acc.x);
              acc.y = fma(a1[q], x1[q].y, acc.y);
            }
            if (a.vec2) {
              acc.x = fma(a2[q], x2[q].x, acc.x);
              acc.y = fma(a2[q], x2[q].y, acc.y);
            }
          }
        }
      }
    }
    // join the sub-groups of the warp (lane bits GS, 2GS, ...), then the warps
#pragma unroll
    for (int off = GS; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    }
    if (grp == 0 && colok) part[wid][gl] = acc;
    __syncthreads();
    const size_t rb = (size_t)row * ld;
    for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
      double2 sres = part[0][c];
#pragma unroll
      for (int w = 1; w < NW; ++w) {
        sres.x += part[w][c].x;
        sres.y += part[w][c].y;
      }
      if (a.beta != 0.0) {
        const double2 o = ld2(a.out + rb + 2 * c);
        sres.x = fma(a.beta, o.x, sres.x);
        sres.y = fma(a.beta, o.y, sres.y);
      }
      st2(a.out + rb + 2 * c, sres);
    }
    __syncthreads();
  }
}

// cost scalars: f = <C,X> + sigma/2 |r|^2  (ManiSDP_unitdiag.m:156)
__global__ void k_cost_finish(RtrState* st, double sigma, int mode, int w, int dual) {
  // dual handles: tmp[6] = sigma/2 |Z'Z|_F^2 + sigma/2 |Af|^2 + k0 (dual.cu: msdp_dual_cost_extra)
  const double f = st->tmp[T_CX] + 0.5 * sigma * st->tmp[T_RR] + (dual ? st->tmp[6] : 0.0);
  st->tmp[T_F] = f;
  if (mode == CG_COSTONLY) return;
  st->cx[w] = st->tmp[T_CX];
  st->rr[w] = st->tmp[T_RR];
  if (mode == CG_INIT) st->fx = f;
  if (mode == CG_TR) {
    st->fprop = f;
    st->gradnorm2_prop = st->gradnorm2;  // the gradient at an accepted proposal is computed afterwards (predicated)
    tr_decide(st);
  }
}

// gradient epilogue.  In: G = EG = 2*eS*Y.  Out: Riemannian gradient, |G|^2 (+ row multipliers / z).
//   oblique : YeG_i = <Y_i, EG_i>, G = EG - Y.*YeG                 (ManiSDP_unitdiag.m:161-163)
//   sphere  : G = EG - <Y,EG> Y  with <Y,EG> = 2z in st->tmp[T_S]   (ManiSDP_unittrace.m:162-163)
//   euclid  : G = EG                                                (ManiSDP.m:157-158)
template <int GS, int VPL, int MF>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_grad_epi(const double* __restrict__ Y, double* __restrict__ G, double* __restrict__ eGout, RtrState* st,
               double* partials, int64_t nrows, int ld, int mode, const int* pred, int zslot) {
  __shared__ double sm[32];
  if (pred && *pred == 0) return;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const double s2z = (MF == MF_SPHERE) ? st->tmp[T_S] : 0.0;
  const long long nob = (MF == MF_OBLIQUE) ? st->nob_rows : 0;
  double q[1] = {0.0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t rb = (size_t)row * ld;
    double2 g[VPL], y[VPL];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        g[t] = ld2(G + rb + 2 * c);
        if (MF != MF_EUCLID) {
          y[t] = ld2(Y + rb + 2 * c);
          dot += y[t].x * g[t].x + y[t].y * g[t].y;
        }
      }
    }
    if (MF == MF_OBLIQUE) {
      dot = rowsel(group_sum<GS>(dot, mask), row < nob);  // Euclidean block of a multi-block point: G = EG (:219-223)
      if (gl == 0) eGout[row] = dot;
    }
    if (MF == MF_SPHERE) dot = s2z;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        if (MF != MF_EUCLID) {
          g[t].x -= y[t].x * dot;
          g[t].y -= y[t].y * dot;
          st2(G + rb + 2 * c, g[t]);
        }
        q[0] += g[t].x * g[t].x + g[t].y * g[t].y;
      }
    }
  }
  double tot[1];
  __syncwarp();
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      st->tmp[T_G2] = tot[0];
      if (mode == CG_INIT || mode == CG_TR) st->gradnorm2 = tot[0];
      if (MF == MF_SPHERE) st->zsph[zslot] = 0.5 * s2z;
    }
  }
}

// Hessian epilogue.  In: H = eH = 2*eS*U + 4*sigma*AyU*Y.  Out: Riemannian Hessian, <U,H>.
//   oblique : H = eH - Y.*sum(Y.*eH) - U.*YeG            (ManiSDP_unitdiag.m:170)
//   sphere  : H = eH - <eH,Y> Y - 2 z U, <eH,Y> in tmp    (ManiSDP_unittrace.m:175-176)
//   euclid  : H = eH                                      (ManiSDP.m:164)
template <int GS, int VPL, int MF>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_hess_epi(const double* __restrict__ Y, const double* __restrict__ U, double* __restrict__ H,
               const double* __restrict__ YeG, RtrState* st, double* partials, int64_t nrows, int ld, int tail_mode,
               int zslot) {
  __shared__ double sm[32];
  if (tail_mode != TAIL_NONE && st->stop != 0) return;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const double shy = (MF == MF_SPHERE) ? st->tmp[T_S] : 0.0;
  const double twoz = (MF == MF_SPHERE) ? 2.0 * st->zsph[zslot] : 0.0;
  const long long nob = (MF == MF_OBLIQUE) ? st->nob_rows : 0;
  double q[1] = {0.0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t rb = (size_t)row * ld;
    double2 hv[VPL], y[VPL], u[VPL];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        hv[t] = ld2(H + rb + 2 * c);
        u[t] = ld2(U + rb + 2 * c);
        if (MF != MF_EUCLID) {
          y[t] = ld2(Y + rb + 2 * c);
          dot += y[t].x * hv[t].x + y[t].y * hv[t].y;
        }
      }
    }
    double mu = 0.0;
    if (MF == MF_OBLIQUE) {
      dot = rowsel(group_sum<GS>(dot, mask), row < nob);
      mu = (row < nob) ? YeG[row] : 0.0;
    }
    if (MF == MF_SPHERE) {
      dot = shy;
      mu = twoz;
    }
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        if (MF != MF_EUCLID) {
          hv[t].x = hv[t].x - y[t].x * dot - u[t].x * mu;
          hv[t].y = hv[t].y - y[t].y * dot - u[t].y * mu;
          st2(H + rb + 2 * c, hv[t]);
        }
        q[0] += u[t].x * hv[t].x + u[t].y * hv[t].y;
      }
    }
  }
  if (tail_mode == TAIL_NONE) return;
  double tot[1];
  __syncwarp();
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      if (tail_mode == TAIL_TCG)
        tcg_after_hv(st, tot[0]);
      else
        st->tmp[0] = tot[0];
    }
  }
}

// guarded dot of two n x ld arrays into st->tmp[slot] (skipped when tCG has stopped)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dot_guard(const double* __restrict__ a, const double* __restrict__ b, RtrState* st, double* partials,
                int64_t nvec, int slot, int skip_if_stopped, const int* pred) {
  __shared__ double sm[32];
  if (pred && *pred == 0) return;
  if (skip_if_stopped && st->stop != 0) return;
  double q[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const double2 x = ld2(a + 2 * i), y = ld2(b + 2 * i);
    q[0] += x.x * y.x + x.y * y.y;
  }
  double tot[1];
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) st->tmp[slot] = tot[0];
  }
}

// KKT: Axb = r + y/sigma ; y <- y - sigma*Axb ; sums |Axb|^2 and b'y_new   (ManiSDP_unitdiag.m:61-64)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dual_update(const double* __restrict__ r, double* __restrict__ y, const double* __restrict__ b, double sigma,
                  int64_t m, int update, RtrState* st, double* partials) {
  __shared__ double sm[64];
  double q[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const double axb = r[k] + y[k] / sigma;
    const double yn = y[k] - sigma * axb;
    if (update) y[k] = yn;
    q[0] += axb * axb;
    q[1] += b[k] * yn;
  }
  double tot[2];
  if (grid_sum_last<2>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      st->tmp[T_AUX] = tot[0];
      st->tmp[T_AUX + 1] = tot[1];
    }
  }
}

// z_i = <Y_i, T_i> -> zout (oblique) and sum_i z_i -> st->tmp[slot]
template <int GS, int VPL>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_rowdot(const double* __restrict__ Y, const double* __restrict__ T, double* __restrict__ zout, RtrState* st,
             double* partials, int64_t nrows, int ld, int slot) {
  __shared__ double sm[32];
  const long long nob = st->nob_rows;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS, nvec = ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  double q[1] = {0.0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t rb = (size_t)row * ld;
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        const double2 a = ld2(Y + rb + 2 * c), b = ld2(T + rb + 2 * c);
        dot += a.x * b.x + a.y * b.y;
      }
    }
    dot = rowsel(group_sum<GS>(dot, mask), row < nob);  // multi-block: z only on the unit-diagonal blocks (:84-88)
    if (gl == 0) {
      if (zout) zout[row] = dot;
      q[0] += dot;
    }
  }
  double tot[1];
  __syncwarp();
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) st->tmp[slot] = tot[0];
  }
}

// out_i -= z_i * V_i  (zdiag) or out -= zs * V
__global__ void __launch_bounds__(MSDP_THREADS)
    k_shift_rows(double* __restrict__ out, const double* __restrict__ V, const double* __restrict__ zdiag, double zs,
                 int64_t nrows, int ld) {
  const int64_t total = nrows * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double z = zdiag ? zdiag[i / ld] : zs;
    out[i] -= z * V[i];
  }
}

// ======================================================================================================================
// host side: setup
// ======================================================================================================================
template <typename T>
static int to_dev(manisdp_handle* h, T** dst, const std::vector<T>& v) {
  CUDA_TRY(h, cudaMalloc((void**)dst, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) CUDA_TRY(h, cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MANISDP_OK;
}

int msdp_affine_setup(manisdp_handle* h, const manisdp_problem* pb) {
  if (!pb->At_jc || !pb->At_ir || !pb->At_pr || !pb->b || !pb->c_pr)
    return msdp_fail(h, MANISDP_E_ARG, "affine kinds need At (CSC), b and c");
  const int64_t n = h->n, m = h->m;
  const bool mb = (h->kind == MANISDP_MULTIBLOCK);
  // rows of At / entries of c: n*n for one block, sum(n_i^2) -- the stacked vecs -- for a multi-block problem
  const uint64_t nn = mb ? (uint64_t)h->mb_off2.back() : (uint64_t)n * (uint64_t)n;
  // r -> (i, j) of the (embedded) matrix, in 64-bit integer arithmetic.  Multi-block: block k = the one whose range
  // [off2_k, off2_k + n_k^2) holds r; inside it r - off2_k = b*n_k + a addresses X_k(a, b) (ManiSDP_multiblock.m:67-72).
  auto split = [&](uint64_t r, int* i, int* j) {
    if (!mb) {
      *i = (int)(r % (uint64_t)n);
      *j = (int)(r / (uint64_t)n);
      return;
    }
    const size_t k = (size_t)(std::upper_bound(h->mb_off2.begin(), h->mb_off2.end(), (int64_t)r) - h->mb_off2.begin()) - 1;
    const uint64_t loc = r - (uint64_t)h->mb_off2[k], nk = (uint64_t)h->mb_n[k];
    *i = (int)((uint64_t)h->mb_roff[k] + loc % nk);
    *j = (int)((uint64_t)h->mb_roff[k] + loc / nk);
  };
  if (n >= (1ll << 31) || m >= (1ll << 31)) return msdp_fail(h, MANISDP_E_ARG, "n and m must be < 2^31");
  const uint64_t nnzA = pb->At_jc[m];
  if (nnzA >= (1ull << 31)) return msdp_fail(h, MANISDP_E_ARG, "nnz(At) must be < 2^31");
  const uint64_t c_nnz = pb->c_ir ? (uint64_t)pb->c_nnz : nn;
  if (!pb->c_ir && (uint64_t)pb->c_nnz != nn) return msdp_fail(h, MANISDP_E_ARG, "dense c must have n*n entries");
  // ---- representation choice
  bool a_dense = (double)nnzA > 0.25 * (double)nn;
  uint64_t c_true_nnz = c_nnz;
  if (!pb->c_ir) {
    c_true_nnz = 0;
    for (uint64_t i = 0; i < nn; ++i) c_true_nnz += (pb->c_pr[i] != 0.0);
  }
  bool s_dense = a_dense || ((double)c_true_nnz + (double)nnzA > 0.125 * (double)nn);
  if (pb->force_mode & 1) s_dense = true;
  if (pb->force_mode & 2) s_dense = false;
  if (pb->force_mode & 4) a_dense = true;
  if (pb->force_mode & 8) a_dense = false;
  if (mb) a_dense = s_dense = false;  // the block structure lives in the index lists; an N x N dense S would not
  if (a_dense) s_dense = true;
  if (s_dense && n > 40000) return msdp_fail(h, MANISDP_E_ARG, "dense S representation needs n <= 40000");
  h->a_mode = a_dense ? MODE_DENSE : MODE_SPARSE;
  h->s_mode = s_dense ? MODE_DENSE : MODE_SPARSE;

  // ---- vectors of length m
  std::vector<double> bh(pb->b, pb->b + m), zeros((size_t)m, 0.0);
  double nb = 0.0;
  for (double v : bh) nb += v * v;
  h->normb = 1.0 + sqrt(nb);
  MSDP_TRY(to_dev(h, &h->b, bh));
  MSDP_TRY(to_dev(h, &h->y, zeros));
  MSDP_TRY(to_dev(h, &h->resid[0], zeros));
  MSDP_TRY(to_dev(h, &h->resid[1], zeros));
  MSDP_TRY(to_dev(h, &h->wU, zeros));
  MSDP_TRY(to_dev(h, &h->wtmp, zeros));

  // ---- split the 64-bit linear indices of At (bit-exact integer arithmetic)
  std::vector<int> ei((size_t)nnzA), ej((size_t)nnzA), kptr((size_t)m + 1);
  std::vector<double> ea(pb->At_pr, pb->At_pr + nnzA);
  for (int64_t k = 0; k <= m; ++k) kptr[(size_t)k] = (int)pb->At_jc[k];
  for (uint64_t e = 0; e < nnzA; ++e) {
    const uint64_t r = pb->At_ir[e];
    if (r >= nn) return msdp_fail(h, MANISDP_E_ARG, "At: row index out of range");
    split(r, &ei[(size_t)e], &ej[(size_t)e]);
  }
  if (!a_dense) {
    ASparse& S = h->As;
    S.nnz = (int64_t)nnzA;
    MSDP_TRY(to_dev(h, &S.kptr, kptr));
    {  // segments of at most SDDMM_SEG entries, built only when some constraint is longer than 2 segments
      int maxlen = 0;
      for (int64_t k = 0; k < m; ++k) maxlen = std::max(maxlen, kptr[(size_t)k + 1] - kptr[(size_t)k]);
      if (maxlen > 2 * SDDMM_SEG) {
        std::vector<int> sptr, ksegs((size_t)m + 1, 0);
        for (int64_t k = 0; k < m; ++k) {
          ksegs[(size_t)k] = (int)sptr.size();
          for (int e = kptr[(size_t)k]; e < kptr[(size_t)k + 1]; e += SDDMM_SEG) sptr.push_back(e);
        }
        ksegs[(size_t)m] = (int)sptr.size();
        S.nseg = (int64_t)sptr.size();
        sptr.push_back(kptr[(size_t)m]);
        // a segment ends where the next one starts, except at a constraint boundary: make that explicit
        std::vector<int> send(sptr);
        MSDP_TRY(to_dev(h, &S.sptr, send));
        MSDP_TRY(to_dev(h, &S.ksegs, ksegs));
        CUDA_TRY(h, cudaMalloc((void**)&S.segval, (size_t)std::max<int64_t>(1, S.nseg) * sizeof(double)));
      }
    }
    MSDP_TRY(to_dev(h, &S.ei, ei));
    MSDP_TRY(to_dev(h, &S.ej, ej));
    MSDP_TRY(to_dev(h, &S.ea, ea));
    // by-row lists (stable counting sort on i keeps k ascending inside a row: deterministic summation order)
    std::vector<int> rptr((size_t)n + 1, 0), rj((size_t)nnzA), rk((size_t)nnzA);
    std::vector<double> ra((size_t)nnzA);
    for (uint64_t e = 0; e < nnzA; ++e) rptr[(size_t)ei[e] + 1]++;
    for (int64_t i = 0; i < n; ++i) rptr[(size_t)i + 1] += rptr[(size_t)i];
    std::vector<int> fill(rptr.begin(), rptr.end() - 1);
    for (int64_t k = 0; k < m; ++k)
      for (int e = kptr[(size_t)k]; e < kptr[(size_t)k + 1]; ++e) {
        const int pos = fill[(size_t)ei[e]]++;
        rj[(size_t)pos] = ej[e];
        rk[(size_t)pos] = (int)k;
        ra[(size_t)pos] = ea[e];
      }
    MSDP_TRY(to_dev(h, &S.rptr, rptr));
    MSDP_TRY(to_dev(h, &S.rj, rj));
    MSDP_TRY(to_dev(h, &S.rk, rk));
    MSDP_TRY(to_dev(h, &S.ra, ra));
  }
  if (s_dense) {
    ADense& D = h->Ad;
    D.nnz = (int64_t)nnzA;
    std::vector<int> klin((size_t)nnzA);
    for (uint64_t e = 0; e < nnzA; ++e) klin[(size_t)e] = (int)pb->At_ir[e];  // n <= 40000 => fits int32
    if (a_dense) {
      MSDP_TRY(to_dev(h, &D.kptr, kptr));
      MSDP_TRY(to_dev(h, &D.klin, klin));
      MSDP_TRY(to_dev(h, &D.ka, ea));
    }
    // transpose over the touched positions (stable: k ascending inside a position)
    std::vector<int> order((size_t)nnzA);
    std::vector<int> kof((size_t)nnzA);
    for (int64_t k = 0; k < m; ++k)
      for (int e = kptr[(size_t)k]; e < kptr[(size_t)k + 1]; ++e) kof[(size_t)e] = (int)k;
    {  // stable counting sort of the entries by linear position (positions < n*n: one pass instead of a comparison sort
       // of 4.8 M entries on BQP-60)
      std::vector<int> cnt((size_t)nn + 1, 0);
      for (uint64_t e = 0; e < nnzA; ++e) cnt[(size_t)klin[(size_t)e] + 1]++;
      for (uint64_t q = 0; q < nn; ++q) cnt[(size_t)q + 1] += cnt[(size_t)q];
      for (uint64_t e = 0; e < nnzA; ++e) order[(size_t)cnt[(size_t)klin[(size_t)e]]++] = (int)e;
    }
    std::vector<int> upos, lptr, lk((size_t)nnzA);
    std::vector<double> la((size_t)nnzA);
    for (size_t q = 0; q < order.size(); ++q) {
      const int e = order[q];
      if (q == 0 || klin[(size_t)e] != klin[(size_t)order[q - 1]]) {
        upos.push_back(klin[(size_t)e]);
        lptr.push_back((int)q);
      }
      lk[q] = kof[(size_t)e];
      la[q] = ea[(size_t)e];
    }
    lptr.push_back((int)nnzA);
    D.nu = (int64_t)upos.size();
    MSDP_TRY(to_dev(h, &D.upos, upos));
    MSDP_TRY(to_dev(h, &D.lptr, lptr));
    MSDP_TRY(to_dev(h, &D.lk, lk));
    MSDP_TRY(to_dev(h, &D.la, la));
    // dense C (column-major vec as given; symmetric)
    std::vector<double> Cd((size_t)nn, 0.0);
    if (pb->c_ir) {
      for (int64_t q = 0; q < pb->c_nnz; ++q) {
        if (pb->c_ir[q] >= nn) return msdp_fail(h, MANISDP_E_ARG, "c: index out of range");
        Cd[(size_t)pb->c_ir[q]] += pb->c_pr[q];
      }
    } else {
      std::copy(pb->c_pr, pb->c_pr + nn, Cd.begin());
    }
    MSDP_TRY(to_dev(h, &h->Cdense, Cd));
    CUDA_TRY(h, cudaMalloc((void**)&h->eS, (size_t)nn * sizeof(double)));
    CUDA_TRY(h, cudaMemcpy(h->eS, h->Cdense, (size_t)nn * sizeof(double), cudaMemcpyDeviceToDevice));
    if (a_dense) {
      CUDA_TRY(h, cudaMalloc((void**)&h->Mbuf, (size_t)nn * sizeof(double)));
      CUDA_TRY(h, cudaMalloc((void**)&h->Tbuf, (size_t)nn * sizeof(double)));
      CUDA_TRY(h, cudaMemset(h->Tbuf, 0, (size_t)nn * sizeof(double)));
    }
  } else {
    // sparse C as row lists: out_i = sum_j C(i,j) V_j
    std::vector<std::pair<uint64_t, double>> ent;
    if (pb->c_ir) {
      for (int64_t q = 0; q < pb->c_nnz; ++q) ent.push_back({pb->c_ir[q], pb->c_pr[q]});
    } else {
      for (uint64_t r = 0; r < nn; ++r)
        if (pb->c_pr[r] != 0.0) ent.push_back({r, pb->c_pr[r]});
    }
    std::vector<int> rp((size_t)n + 1, 0), col(ent.size());
    std::vector<double> val(ent.size());
    for (auto& t : ent) {
      if (t.first >= nn) return msdp_fail(h, MANISDP_E_ARG, "c: index out of range");
      int ci, cj;
      split(t.first, &ci, &cj);
      rp[(size_t)ci + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) rp[(size_t)i + 1] += rp[(size_t)i];
    std::vector<int> fill(rp.begin(), rp.end() - 1);
    for (auto& t : ent) {
      int i, j;
      split(t.first, &i, &j);
      const int pos = fill[(size_t)i]++;
      col[(size_t)pos] = j;
      val[(size_t)pos] = t.second;
    }
    h->C.nrows = n;
    h->C.nnz = (int64_t)ent.size();
    MSDP_TRY(to_dev(h, &h->C.rowptr, rp));
    MSDP_TRY(to_dev(h, &h->C.col, col));
    MSDP_TRY(to_dev(h, &h->C.val, val));
  }
  return MANISDP_OK;
}

void msdp_affine_free(manisdp_handle* h) {
  void* ptrs[] = {h->Cdense, h->eS,     h->Mbuf,   h->Tbuf,   h->b,       h->y,       h->resid[0], h->resid[1],
                  h->wU,     h->wtmp,   h->As.kptr, h->As.ei, h->As.ej,   h->As.ea,   h->As.rptr,  h->As.rj,
                  h->As.rk,  h->As.ra,  h->As.sptr, h->As.ksegs, h->As.segval, h->Ad.kptr, h->Ad.klin, h->Ad.ka, h->Ad.upos, h->Ad.lptr,  h->Ad.lk,
                  h->Ad.la};
  for (void* p : ptrs)
    if (p) cudaFree(p);
}

// ======================================================================================================================
// host side: closures
// ======================================================================================================================
static const int* accepted_flag(manisdp_handle* h) { return &h->st->accepted; }

// w = A(P Q') into `out`; mode 1 turns it into the residual r and leaves |r|^2 in tmp[T_RR]
static int apply_A(manisdp_handle* h, const double* P, const double* Q, double* out, int mode, int skip_if_stopped) {
  const int ld = (int)h->ld;
  const double inv_sigma = h->dual.on ? 0.0 : 1.0 / h->sigma;  // dual handles: r = A~ vec(S) - A~ c, h->y is the ADMM's y
  if (h->a_mode == MODE_DENSE) {
    MSDP_TRY(msdp_gemm_nt(h, P, ld, Q, ld, (int)h->n, ld, h->Mbuf, 1.0, skip_if_stopped ? &h->st->stop : nullptr, 1));
    k_gather_spmv<<<mgrid(h, h->m), MSDP_THREADS, 0, h->stream>>>(h->Ad.kptr, h->Ad.klin, h->Ad.ka, h->Mbuf, h->m, out,
                                                                mode, h->b, h->y, inv_sigma, h->st, h->partials,
                                                                skip_if_stopped);
  } else {
    if (h->As.nseg > 0) {  // some constraint is long (theta: the trace row): segments, then the per-constraint sums
      DISPATCH_GEOM(row_geom(h->ld), {
        k_sddmm_seg<GS, VPL><<<mgrid(h, h->As.nseg, MSDP_THREADS / GS), MSDP_THREADS, 0, h->stream>>>(
            h->As.sptr, h->As.ei, h->As.ej, h->As.ea, P, Q, ld, h->As.nseg, h->As.segval, h->st, skip_if_stopped);
      });
      KERNEL_CHECK(h);
      k_sddmm_finish<<<mgrid(h, h->m), MSDP_THREADS, 0, h->stream>>>(h->As.ksegs, h->As.segval, h->m, out, mode, h->b,
                                                                   h->y, inv_sigma, h->st, h->partials, skip_if_stopped);
    } else {
      DISPATCH_GEOM(row_geom(h->ld), {
        k_sddmm<GS, VPL><<<mgrid(h, h->m, MSDP_THREADS / GS), MSDP_THREADS, 0, h->stream>>>(
            h->As.kptr, h->As.ei, h->As.ej, h->As.ea, P, Q, ld, h->m, out, mode, h->b, h->y, inv_sigma, h->st,
            h->partials, skip_if_stopped);
      });
    }
  }
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// out (n x ld) = alpha * (C + coef * At(vec)) * V   using the dense eS buffer (must already hold that matrix) or the
// sparse rowlist; optional second rowlist term c2 * sum a vec2_k V2_j; beta accumulates into out
static int apply_S_sparse(manisdp_handle* h, const double* V1, const double* vec1, double c1, const double* V2,
                          const double* vec2, double c2, double alphaC, double* out, double beta, int ld,
                          const int* pred, int skip_if_stopped, bool with_C) {
  RowlistArgs a{};
  if (with_C) {
    a.crowptr = h->C.rowptr;
    a.ccol = h->C.col;
    a.cval = h->C.val;
  }
  a.alphaC = alphaC;
  a.rptr = h->As.rptr;
  a.rj = h->As.rj;
  a.rk = h->As.rk;
  a.ra = h->As.ra;
  a.vec1 = vec1;
  a.vec2 = vec2;
  a.c1 = c1;
  a.c2 = c2;
  a.V1 = V1;
  a.V2 = V2;
  a.out = out;
  a.beta = beta;
  a.nrows = h->n;
  a.ld = ld;
  a.pred = pred;
  a.skip_if_stopped = skip_if_stopped;
  // long rows (on average >= 64 entries) and full-warp row groups: one CTA per row
  static const int wide_on = getenv("MANISDP_K3_WIDE") ? atoi(getenv("MANISDP_K3_WIDE")) : 1;  // A/B switch
  if (wide_on && h->As.nnz >= 64 * h->n && ld <= MSDP_MAX_LD && h->n <= (int64_t)h->num_sms * 64) {
    const int nb = (int)std::min<int64_t>(h->n, (int64_t)h->num_sms * 8);
    DISPATCH_GEOM(row_geom(ld), {
      if (GS == 32 && VPL <= 8)
        k_rowlist_apply_wide<(VPL <= 8 ? VPL : 8)><<<nb, MSDP_THREADS, 0, h->stream>>>(a, h->st);
      else if (GS < 32)
        k_rowlist_apply_wide_narrow<(GS < 32 ? GS : 16)><<<nb, MSDP_THREADS, 0, h->stream>>>(a, h->st);
    });
    KERNEL_CHECK(h);
    return MANISDP_OK;
  }
  DISPATCH_GEOM(row_geom(ld), {
    k_rowlist_apply<GS, VPL><<<rows_grid(h, h->n, GS), MSDP_THREADS, 0, h->stream>>>(a, h->st);
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// eS <- C + coef * mat(At * vec)   (dense S)
static int form_eS(manisdp_handle* h, const double* vec, double coef, const int* pred) {
  k_touch_update<<<mgrid(h, h->Ad.nu), MSDP_THREADS, 0, h->stream>>>(h->Ad.upos, h->Ad.lptr, h->Ad.lk, h->Ad.la, vec,
                                                                    coef, h->Cdense, h->eS, h->Ad.nu, pred, h->st, 0);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

static int cost_at(manisdp_handle* h, const double* Z, double* resid_out, int mode, int w) {
  const int64_t nn = h->n * h->n;
  // r = A(ZZ') - b - y/sigma, |r|^2
  MSDP_TRY(apply_A(h, Z, Z, resid_out, 1, 0));
  // <C, ZZ'>
  if (h->a_mode == MODE_DENSE) {  // X = ZZ' is in Mbuf
    k_dot_len<<<mgrid(h, nn), MSDP_THREADS, 0, h->stream>>>(h->Cdense, h->Mbuf, nn, h->st, h->partials, T_CX);
    KERNEL_CHECK(h);
  } else if (h->s_mode == MODE_DENSE) {  // sum((C*Z).*Z): one GEMM into Hslot-free scratch (Hd is free outside tCG)
    MSDP_TRY(msdp_gemm_nn(h, h->Cdense, (int)h->n, Z, (int)h->ld, (int)h->ld, h->Hd, (int)h->ld, 1.0, 0.0, nullptr));
    k_dot_guard<<<mgrid(h, h->n * h->ld / 2), MSDP_THREADS, 0, h->stream>>>(h->Hd, Z, h->st, h->partials,
                                                                         h->n * h->ld / 2, T_CX, 0, nullptr);
    KERNEL_CHECK(h);
  } else {
    MSDP_TRY(apply_S_sparse(h, Z, nullptr, 0.0, nullptr, nullptr, 0.0, 1.0, h->Hd, 0.0, (int)h->ld, nullptr, 0, true));
    k_dot_guard<<<mgrid(h, h->n * h->ld / 2), MSDP_THREADS, 0, h->stream>>>(h->Hd, Z, h->st, h->partials,
                                                                         h->n * h->ld / 2, T_CX, 0, nullptr);
    KERNEL_CHECK(h);
  }
  if (h->dual.on) MSDP_TRY(msdp_dual_cost_extra(h, Z, resid_out, mode == CG_COSTONLY ? -1 : w));
  k_cost_finish<<<1, 1, 0, h->stream>>>(h->st, csig(h), mode, w, h->dual.on);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

static int grad_at(manisdp_handle* h, int w, int mode, const int* pred) {
  const double* Z = h->Ybuf[w];
  double* G = h->Gbuf[w];
  const int ld = (int)h->ld;
  const int64_t nvec = h->n * h->ld / 2;
  if (h->s_mode == MODE_DENSE) {
    MSDP_TRY(form_eS(h, h->resid[w], csig(h), pred));  // eS = C + sigma*At*Axb   (ManiSDP_unitdiag.m:160)
    MSDP_TRY(msdp_gemm_nn(h, h->eS, (int)h->n, Z, ld, ld, G, ld, 2.0, 0.0, pred));  // eG = 2*Y*eS (:161)
  } else {
    MSDP_TRY(apply_S_sparse(h, Z, h->resid[w], 2.0 * h->sigma, nullptr, nullptr, 0.0, 2.0, G, 0.0, ld, pred, 0, true));
  }
  if (h->dual.on) MSDP_TRY(msdp_dual_grad_extra(h, Z, G, w, pred));  // + 2*sigma*Y*(Y'Y): X = eS + sigma*S (ManiDSDP_unitdiag.m:181-182)
  if (h->mf == MF_SPHERE) {
    k_dot_guard<<<mgrid(h, nvec), MSDP_THREADS, 0, h->stream>>>(Z, G, h->st, h->partials, nvec, T_S, 0, pred);
    KERNEL_CHECK(h);
  }
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->n, GS);
    switch (h->mf) {
      case MF_OBLIQUE:
        k_grad_epi<GS, VPL, MF_OBLIQUE><<<nb, MSDP_THREADS, 0, h->stream>>>(Z, G, h->eG[w], h->st, h->partials, h->n,
                                                                         ld, mode, pred, 0);
        break;
      case MF_SPHERE:
        k_grad_epi<GS, VPL, MF_SPHERE><<<nb, MSDP_THREADS, 0, h->stream>>>(Z, G, h->eG[w], h->st, h->partials, h->n,
                                                                        ld, mode, pred, 0);
        break;
      default:
        k_grad_epi<GS, VPL, MF_EUCLID><<<nb, MSDP_THREADS, 0, h->stream>>>(Z, G, h->eG[w], h->st, h->partials, h->n,
                                                                        ld, mode, pred, 0);
    }
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// which >= 0 explicit buffer; -1 proposal (h->pt ^ 1); -2 current (h->pt).  The host mirror h->pt is exact here: the
// driver reads the state back once per TR iteration (rtr.cu) and graphs are built per value of pt.
int msdp_affine_costgrad(manisdp_handle* h, int which, int cg_mode) {
  const int w = which >= 0 ? which : (which == -1 ? (h->pt ^ 1) : h->pt);
  if (h->dual.on && h->dual.dirty) MSDP_TRY(msdp_dual_refresh(h));
  if (cg_mode == CG_TR_DEFER) return msdp_fail(h, MANISDP_E_ARG, "affine kinds are not row-sharded");
  MSDP_TRY(cost_at(h, h->Ybuf[w], h->resid[w], cg_mode, w));
  if (cg_mode == CG_COSTONLY) return MANISDP_OK;
  // the gradient at a proposal is only needed (and only allowed to overwrite eS) when the step was accepted
  const int* pred = (cg_mode == CG_TR) ? accepted_flag(h) : nullptr;
  return grad_at(h, w, cg_mode, pred);
}

int msdp_affine_hess(manisdp_handle* h, const double* D, double* Hout, int tail_mode) {
  const int w = h->pt;
  const double* Y = h->Ybuf[w];
  const int ld = (int)h->ld;
  const int skip = (tail_mode != TAIL_NONE);
  const int64_t nvec = h->n * h->ld / 2;
  const double s4 = 4.0 * csig(h);
  // AyU-part: wU = A(U Y')  (ManiSDP_unitdiag.m:167-168 / ManiSDP.m:162-163)
  MSDP_TRY(apply_A(h, D, Y, h->wU, 0, skip));
  if (h->s_mode == MODE_DENSE) {
    const int* stopf = skip ? &h->st->stop : nullptr;
    MSDP_TRY(msdp_gemm_nn(h, h->eS, (int)h->n, D, ld, ld, Hout, ld, 2.0, 0.0, stopf, 1));  // 2*U*eS
    if (h->a_mode == MODE_DENSE) {
      k_touch_update<<<mgrid(h, h->Ad.nu), MSDP_THREADS, 0, h->stream>>>(h->Ad.upos, h->Ad.lptr, h->Ad.lk, h->Ad.la,
                                                                        h->wU, 1.0, nullptr, h->Tbuf, h->Ad.nu,
                                                                        nullptr, h->st, skip);
      KERNEL_CHECK(h);
      MSDP_TRY(msdp_gemm_nn(h, h->Tbuf, (int)h->n, Y, ld, ld, Hout, ld, s4, 1.0, stopf, 1));  // + 4*sigma*Y*AyU
    } else {
      MSDP_TRY(apply_S_sparse(h, nullptr, nullptr, 0.0, Y, h->wU, s4, 0.0, Hout, 1.0, ld, nullptr, skip, false));
    }
  } else {
    // 2*(C + sigma*At(r))*U + 4*sigma*At(wU)*Y in one rowlist pass
    MSDP_TRY(apply_S_sparse(h, D, h->resid[w], 2.0 * h->sigma, Y, h->wU, s4, 2.0, Hout, 0.0, ld, nullptr, skip, true));
  }
  // dual handles: + 2*sigma*(Y*(Y'U + U'Y) + U*(Y'Y))   (ManiDSDP_unitdiag.m:189 with X = eS + sigma*S)
  if (h->dual.on) MSDP_TRY(msdp_dual_hess_extra(h, Y, D, Hout, w, skip));
  if (h->mf == MF_SPHERE) {
    k_dot_guard<<<mgrid(h, nvec), MSDP_THREADS, 0, h->stream>>>(Hout, Y, h->st, h->partials, nvec, T_S, skip, nullptr);
    KERNEL_CHECK(h);
  }
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->n, GS);
    switch (h->mf) {
      case MF_OBLIQUE:
        k_hess_epi<GS, VPL, MF_OBLIQUE><<<nb, MSDP_THREADS, 0, h->stream>>>(Y, D, Hout, h->eG[w], h->st, h->partials,
                                                                         h->n, ld, tail_mode, 0);
        break;
      case MF_SPHERE:
        k_hess_epi<GS, VPL, MF_SPHERE><<<nb, MSDP_THREADS, 0, h->stream>>>(Y, D, Hout, h->eG[w], h->st, h->partials,
                                                                        h->n, ld, tail_mode, 0);
        break;
      default:
        k_hess_epi<GS, VPL, MF_EUCLID><<<nb, MSDP_THREADS, 0, h->stream>>>(Y, D, Hout, h->eG[w], h->st, h->partials,
                                                                        h->n, ld, tail_mode, 0);
    }
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_affine_cost_only(manisdp_handle* h, const double* Z, double* f_host) {
  if (h->dual.on && h->dual.dirty) MSDP_TRY(msdp_dual_refresh(h));
  MSDP_TRY(cost_at(h, Z, h->wtmp, CG_COSTONLY, 0));
  const int keep = h->pt;
  MSDP_TRY(msdp_sync_state(h));
  h->pt = keep;
  if (f_host) *f_host = h->st_host->tmp[T_F];
  return MANISDP_OK;
}

// ======================================================================================================================
// KKT step and dual-slack operator
// ======================================================================================================================
int msdp_affine_kkt(manisdp_handle* h, int update_dual, manisdp_kkt_info* out) {
  // Axb = A(X) - b is recovered from the cached residual r = Axb - y/sigma of the current point
  h->cache_valid = 0;
  MSDP_TRY(msdp_ensure_costgrad(h));
  const int w = h->pt;
  const double* Y = h->Ybuf[w];
  const int ld = (int)h->ld;
  double* ykkt = h->y;
  if (!update_dual) {  // evaluate with a scratch copy of y
    CUDA_TRY(h, cudaMemcpyAsync(h->wtmp, h->y, (size_t)h->m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    ykkt = h->wtmp;
  }
  k_dual_update<<<mgrid(h, h->m), MSDP_THREADS, 0, h->stream>>>(h->resid[w], ykkt, h->b, h->sigma, h->m, 1, h->st,
                                                              h->partials);
  KERNEL_CHECK(h);
  // dual slack eS = C - At*y (ManiSDP_unitdiag.m:65) and z
  double* T = h->Hd;  // scratch n x ld
  if (h->s_mode == MODE_DENSE) {
    MSDP_TRY(form_eS(h, ykkt, -1.0, nullptr));
    if (h->mf != MF_EUCLID) MSDP_TRY(msdp_gemm_nn(h, h->eS, (int)h->n, Y, ld, ld, T, ld, 1.0, 0.0, nullptr));
  } else if (h->mf != MF_EUCLID) {
    MSDP_TRY(apply_S_sparse(h, Y, ykkt, -1.0, nullptr, nullptr, 0.0, 1.0, T, 0.0, ld, nullptr, 0, true));
  }
  if (h->mf != MF_EUCLID) {
    DISPATCH_GEOM(row_geom(h->ld), {
      k_rowdot<GS, VPL><<<rows_grid(h, h->n, GS), MSDP_THREADS, 0, h->stream>>>(
          Y, T, h->mf == MF_OBLIQUE ? h->zdiag : nullptr, h->st, h->partials, h->n, ld, T_S);
    });
    KERNEL_CHECK(h);
  }
  const int keep = h->pt;
  MSDP_TRY(msdp_sync_state(h));
  h->pt = keep;
  const RtrState* s = h->st_host;
  const double obj = s->cx[w];
  const double pinf = sqrt(s->tmp[T_AUX]) / h->normb;
  double by = s->tmp[T_AUX + 1];
  double zsum = 0.0;
  if (h->mf != MF_EUCLID) {
    zsum = s->tmp[T_S];
    by += zsum;  // by = b'y + sum(z) (unitdiag :70) / + z (unittrace :70)
  }
  h->zshift = (h->mf == MF_SPHERE) ? zsum : 0.0;
  h->y_kkt = ykkt;
  out->obj = obj;
  out->by = by;
  out->pinf = pinf;
  out->z_sum = zsum;
  out->gap = fabs(obj - by) / (fabs(by) + fabs(obj) + 1.0);  // :71
  // y changed (or eS now holds the KKT slack): the closure caches of the point are stale
  h->cache_valid = 0;
  h->grad_valid = 0;
  return MANISDP_OK;
}

// helpers shared with dual.cu: z_i = <Y_i, T_i> (+ sum into st->tmp[slot]); dst[touched] = base + coef * mat(At * vec)
int msdp_affine_rowdot(manisdp_handle* h, const double* Y, const double* T, double* zout, int slot) {
  const int ld = (int)h->ld;
  DISPATCH_GEOM(row_geom(h->ld), {
    k_rowdot<GS, VPL><<<rows_grid(h, h->n, GS), MSDP_THREADS, 0, h->stream>>>(Y, T, zout, h->st, h->partials, h->n, ld,
                                                                             slot);
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}
int msdp_affine_touch(manisdp_handle* h, const double* vec, double coef, const double* base, double* dst) {
  k_touch_update<<<mgrid(h, h->Ad.nu), MSDP_THREADS, 0, h->stream>>>(h->Ad.upos, h->Ad.lptr, h->Ad.lk, h->Ad.la, vec,
                                                                    coef, base, dst, h->Ad.nu, nullptr, h->st, 0);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// AV = S*V with S = eS - diag(z) (oblique) / eS - z*I (sphere) / eS (Euclidean), eS = C - At*y of the last kkt call
int msdp_affine_apply_S(manisdp_handle* h, const double* V, double* AV, int kld) {
  if (h->s_mode == MODE_DENSE) {
    MSDP_TRY(msdp_gemm_nn(h, h->eS, (int)h->n, V, kld, kld, AV, kld, 1.0, 0.0, nullptr));
  } else {
    MSDP_TRY(apply_S_sparse(h, V, h->y_kkt ? h->y_kkt : h->y, -1.0, nullptr, nullptr, 0.0, 1.0, AV, 0.0, kld, nullptr, 0, true));
  }
  if (h->mf != MF_EUCLID) {
    k_shift_rows<<<mgrid(h, h->n * kld), MSDP_THREADS, 0, h->stream>>>(AV, V, h->mf == MF_OBLIQUE ? h->zdiag : nullptr,
                                                                     h->zshift, h->n, kld);
    KERNEL_CHECK(h);
  }
  return MANISDP_OK;
}

// Read-back of the (i, j) index split of At's rows (tests: "A(YY') index handling must be bit-exact").  Sparse A: the
// int32 arrays the SDDMM / row-list kernels index with; dense A: the int32 linear positions the gather kernel indexes
// the n x n matrix with, split here with the same 64-bit integer div / mod used at create.  CSC order of At.
int msdp_affine_index_split(manisdp_handle* h, int64_t* i_out, int64_t* j_out, int64_t cap, int64_t* count) {
  const int64_t nnz = (h->a_mode == MODE_DENSE) ? h->Ad.nnz : h->As.nnz;
  if (count) *count = nnz;
  if (!i_out || !j_out) return MANISDP_OK;
  const int64_t k = std::min<int64_t>(cap, nnz);
  if (k <= 0) return MANISDP_OK;
  std::vector<int> a((size_t)k), b((size_t)k);
  if (h->a_mode == MODE_DENSE) {
    CUDA_TRY(h, cudaMemcpy(a.data(), h->Ad.klin, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < k; ++e) {
      const uint64_t r = (uint64_t)(unsigned)a[(size_t)e];
      i_out[e] = (int64_t)(r % (uint64_t)h->n);
      j_out[e] = (int64_t)(r / (uint64_t)h->n);
    }
  } else {
    CUDA_TRY(h, cudaMemcpy(a.data(), h->As.ei, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemcpy(b.data(), h->As.ej, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t e = 0; e < k; ++e) {
      i_out[e] = a[(size_t)e];
      j_out[e] = b[(size_t)e];
    }
  }
  return MANISDP_OK;
}
