// jacobi.cu -- batched dense symmetric eigen-decomposition on the device: one CTA per matrix, cyclic Jacobi with the
// round-robin (tournament) ordering, so that the n/2 rotations of a round touch disjoint row / column pairs and are applied
// together.  An ALTERNATIVE (MANISDP_MB_EIG=device) for the place where the reference calls eig() on many small matrices at
// once -- the dual slack blocks S{i} of a multi-block SDP (ManiSDP_multiblock.m:90, t blocks of order 10..211 here): the
// matrices, which are produced on the device, never travel to the host and all blocks are decomposed concurrently, one SM
// each.  Measured slower than Householder + QL on 16 host threads (multiblock.cu: manisdp_mb_kkt), so not the default.
//
// Per round: (1) one thread per pair computes (c, s) from a_pp, a_qq, a_pq of the matrix as it stands at the start of the
// round; (2) column phase A <- A J, V <- V J; (3) row phase A <- J' A.  J is the product of the round's disjoint plane
// rotations, so the simultaneous application equals the sequential one.  A sweep is n - 1 rounds (n even; an odd order
// gets a phantom index whose pairs are skipped).  Sweeps stop when the off-diagonal Frobenius norm falls below
// 1e-14 |A|_F (quadratic convergence: 6..10 sweeps).  Eigenvalues are returned in ascending order with the eigenvectors
// permuted accordingly (rank by counting, no comparison network).  Matrices live in global memory (L2-resident: 2 n^2
// doubles per CTA); everything is deterministic.
#include <math.h>
#include <algorithm>
#include "common.cuh"

#define JAC_THREADS 512
#define JAC_MAX_N 1024

struct JacobiBatch {
  double* A;            // matrices, row-major, matrix b at A + off[b]; overwritten: on exit column k = k-th eigenvector
  double* V;            // scratch of the same layout
  double* w;            // eigenvalues, matrix b at w + woff[b], ascending
  const int64_t* off;   // element offsets of the matrices
  const int* woff;      // offsets of the eigenvalue vectors
  const int* n;         // orders
  int* sweeps;          // per matrix: sweeps used (negative: not converged within max_sweeps)
  int max_sweeps;
};

__global__ void __launch_bounds__(JAC_THREADS) k_jacobi_batched(JacobiBatch jb) {
  __shared__ int sp[JAC_MAX_N / 2], sq[JAC_MAX_N / 2];
  __shared__ double sc[JAC_MAX_N / 2], ss[JAC_MAX_N / 2];
  __shared__ double red[JAC_THREADS / 32 * 2];
  __shared__ double s_off2, s_tot2;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int n = jb.n[b];
  double* __restrict__ A = jb.A + jb.off[b];
  double* __restrict__ V = jb.V + jb.off[b];
  double* __restrict__ w = jb.w + jb.woff[b];
  for (int e = tid; e < n * n; e += nt) V[e] = (e / n == e % n) ? 1.0 : 0.0;
  const int m = n + (n & 1);  // players of the tournament (phantom index n when n is odd)
  const int half = m / 2;
  int used = 0;
  bool done = (n <= 1);
  __syncthreads();
  for (int sweep = 0; sweep < jb.max_sweeps && !done; ++sweep) {
    for (int r = 0; r < m - 1; ++r) {
      // (1) pairs of the round and their rotations
      for (int k = tid; k < half; k += nt) {
        int p, q;
        if (k == 0) {
          p = m - 1;
          q = r;
        } else {
          p = (r + k) % (m - 1);
          q = (r - k + (m - 1)) % (m - 1);
        }
        if (p > q) {
          const int t = p;
          p = q;
          q = t;
        }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[(size_t)p * n + q];
          if (apq != 0.0) {
            const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            s = t * c;
          }
        } else {
          q = -1;  // pair with the phantom index: nothing to rotate
        }
        sp[k] = p;
        sq[k] = q;
        sc[k] = c;
        ss[k] = s;
      }
      __syncthreads();
      // (2) columns of A and V:  [x_p, x_q] <- [c x_p - s x_q, s x_p + c x_q]
      for (int e = tid; e < n * half; e += nt) {
        const int i = e / half, k = e - i * half;
        const int q = sq[k];
        if (q < 0) continue;
        const int p = sp[k];
        const double c = sc[k], s = ss[k];
        if (s == 0.0) continue;
        const size_t ip = (size_t)i * n + p, iq = (size_t)i * n + q;
        const double ap = A[ip], aq = A[iq];
        A[ip] = c * ap - s * aq;
        A[iq] = s * ap + c * aq;
        const double vp = V[ip], vq = V[iq];
        V[ip] = c * vp - s * vq;
        V[iq] = s * vp + c * vq;
      }
      __syncthreads();
      // (3) rows of A
      for (int e = tid; e < half * n; e += nt) {
        const int k = e / n, j = e - k * n;
        const int q = sq[k];
        if (q < 0) continue;
        const double c = sc[k], s = ss[k];
        if (s == 0.0) continue;
        const size_t pj = (size_t)sp[k] * n + j, qj = (size_t)q * n + j;
        const double ap = A[pj], aq = A[qj];
        A[pj] = c * ap - s * aq;
        A[qj] = s * ap + c * aq;
      }
      __syncthreads();
    }
    used = sweep + 1;
    // off-diagonal and total squared norms (fixed reduction tree)
    double o2 = 0.0, t2 = 0.0;
    for (int e = tid; e < n * n; e += nt) {
      const double v = A[e];
      t2 += v * v;
      if (e / n != e % n) o2 += v * v;
    }
    for (int d = 16; d > 0; d >>= 1) {
      o2 += __shfl_xor_sync(0xffffffffu, o2, d);
      t2 += __shfl_xor_sync(0xffffffffu, t2, d);
    }
    if ((tid & 31) == 0) {
      red[(tid >> 5) * 2] = o2;
      red[(tid >> 5) * 2 + 1] = t2;
    }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, c = 0.0;
      for (int g = 0; g < nt / 32; ++g) {
        a += red[2 * g];
        c += red[2 * g + 1];
      }
      s_off2 = a;
      s_tot2 = c;
    }
    __syncthreads();
    done = (s_off2 <= 1e-28 * s_tot2) || (s_tot2 == 0.0);
    __syncthreads();
  }
  if (tid == 0) jb.sweeps[b] = done ? used : -used;
  // ascending order: rank of eigenvalue i = number of eigenvalues that sort before it (ties by index); the eigenvectors go
  // into the A buffer with their columns permuted accordingly
  __shared__ int srank[JAC_MAX_N];
  for (int i = tid; i < n; i += nt) {
    const double wi = A[(size_t)i * n + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double wj = A[(size_t)j * n + j];
      rank += (wj < wi) || (wj == wi && j < i);
    }
    srank[i] = rank;
    w[rank] = wi;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += nt) {
    const int r = e / n, i = e - r * n;
    A[(size_t)r * n + srank[i]] = V[e];
  }
}

// host launcher: `count` matrices described by device arrays (offsets / orders) already in place
int msdp_jacobi_batched(manisdp_handle* h, double* A, double* V, double* w, const int64_t* off_dev, const int* woff_dev,
                        const int* n_dev, int* sweeps_dev, int count, int max_n) {
  if (max_n > JAC_MAX_N) return msdp_fail(h, MANISDP_E_ARG, "batched Jacobi: matrix order above 1024");
  JacobiBatch jb{A, V, w, off_dev, woff_dev, n_dev, sweeps_dev, 30};
  k_jacobi_batched<<<count, JAC_THREADS, 0, h->stream>>>(jb);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}
