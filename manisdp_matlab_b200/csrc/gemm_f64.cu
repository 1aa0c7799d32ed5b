// gemm_f64.cu -- K4: FP64 tensor-core GEMM for the dense path (BQP / quartic-sphere / theta with dense C).
// tcgen05 has no FP64 kind, so FP64 tensor work on sm_100a is the DMMA path: mma.sync.aligned.m8n8k4.f64.
// One kernel covers the three shapes of the dense closures (ManiSDP_unitdiag.m:160-169, ManiSDP.m:157-164):
//     NN   out(n x w)  = alpha * S(n x n) * V(n x w) + beta * out        (2*eS*U, 4*sigma*AyU*Y, S*V of the eigen step)
//     NT   M(n x n)    = P(n x w) * Q(n x w)'                            (Y'*U of the reference, row layout)
// through runtime strides.  Block tile 64x64x16, 8 warps (4 x 2), warp tile 16x32 = 2x4 DMMA tiles.
#include "gemm.h"

#define BM 64
#define BN 64
#define BK 16
#define PAD 4

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct GemmArgs {
  const double* A;  // M x K : A(m,k) = A[m*sam + k*sak]
  const double* B;  // K x N : B(k,n) = B[k*sbk + n*sbn]
  double* C;        // M x N : C(m,n) = C[m*ldc + n]
  int64_t sam, sak, sbk, sbn, ldc;
  int M, N, K;
  double alpha, beta;
  const int* pred;  // optional device flag
  int pred_sense;   // 0: skip when *pred == 0 ; 1: skip when *pred != 0
};

__global__ void __launch_bounds__(256) k_gemm_f64(GemmArgs g) {
  if (g.pred && ((*g.pred == 0) != (g.pred_sense != 0))) return;
  __shared__ double As[BM][BK + PAD];
  __shared__ double Bs[BK][BN + PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const bool a_kfast = (g.sak == 1);
  const bool b_nfast = (g.sbn == 1);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    // ---- global -> shared (zero fill outside the matrix)
    for (int i = tid; i < BM * BK; i += 256) {
      int m, k;
      if (a_kfast) {
        m = i / BK;
        k = i % BK;
      } else {
        k = i / BM;
        m = i % BM;
      }
      const int gm = m0 + m, gk = k0 + k;
      As[m][k] = (gm < g.M && gk < g.K) ? g.A[(int64_t)gm * g.sam + (int64_t)gk * g.sak] : 0.0;
    }
    for (int i = tid; i < BK * BN; i += 256) {
      int k, n;
      if (b_nfast) {
        k = i / BN;
        n = i % BN;
      } else {
        n = i / BK;
        k = i % BK;
      }
      const int gk = k0 + k, gn = n0 + n;
      Bs[k][n] = (gk < g.K && gn < g.N) ? g.B[(int64_t)gk * g.sbk + (int64_t)gn * g.sbn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = As[wm * 16 + i * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk + (lane & 3)][wn * 32 + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // ---- epilogue: C = alpha*acc + beta*C
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + wm * 16 + i * 8 + (lane >> 2);
      const int n = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
      if (m < g.M) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (n + c < g.N) {
            double* p = g.C + (int64_t)m * g.ldc + n + c;
            const double v = g.alpha * acc[i][j][c];
            *p = (g.beta == 0.0) ? v : (v + g.beta * *p);
          }
        }
      }
    }
}

static int launch(manisdp_handle* h, const GemmArgs& g) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  k_gemm_f64<<<grid, 256, 0, h->stream>>>(g);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// out(n x w, ld = ldo) = alpha * S(n x n, row stride n) * V(n x w, ld = ldv) + beta * out
int msdp_gemm_nn(manisdp_handle* h, const double* S, int n, const double* V, int ldv, int w, double* out, int ldo,
                 double alpha, double beta, const int* pred, int pred_sense) {
  GemmArgs g{S, V, out, n, 1, ldv, 1, ldo, n, w, n, alpha, beta, pred, pred_sense};
  return launch(h, g);
}

// M(n x n, row stride n) = alpha * P(n x w) * Q(n x w)'
int msdp_gemm_nt(manisdp_handle* h, const double* P, int ldp, const double* Q, int ldq, int n, int w, double* M,
                 double alpha, const int* pred, int pred_sense) {
  GemmArgs g{P, Q, M, ldp, 1, 1, ldq, n, n, n, w, alpha, 0.0, pred, pred_sense};
  return launch(h, g);
}
