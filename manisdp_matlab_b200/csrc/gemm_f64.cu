// gemm_f64.cu -- K4: FP64 tensor-core GEMM for the dense path (BQP / quartic-sphere / theta with dense C).
// tcgen05 has no FP64 kind, so FP64 tensor work on sm_100a is the DMMA path: mma.sync.aligned.m8n8k4.f64.
// One kernel covers the shapes of the dense closures (ManiSDP_unitdiag.m:160-169, ManiSDP.m:157-164):
//     NN   out(n x w)  = alpha * S(n x n) * V(n x w) + beta * out        (2*eS*U, 4*sigma*AyU*Y, S*V of the eigen step)
//     NT   M(n x n)    = P(n x w) * Q(n x w)'                            (Y'*U of the reference, row layout)
// through runtime strides.  Block tile 64 x BN x 16 with BN = 64 / 32 / 16 chosen from the output width (the LOBPCG block
// of the eigen step has 12 columns: a 64-wide tile would spend 81 % of its DMMA work on padding), 8 warps (4 x 2), warp
// tile 16 x BN/2 = 2 x NT DMMA tiles, operands staged through a double-buffered shared-memory ring.
//
// SPLIT-K: the NN products are tall-skinny (n = 1831, w = 8..400 on BQP-60): 29 x 5 output tiles at most, each with a
// 115-step K loop, i.e. the launch is latency bound at < 1 CTA per SM (measured 1.0 ms per Hessian product independent
// of w, profiles/r1_dense_hv_bqp60_before_splitk.jsonl).  The K range is therefore cut into S slices so that about
// 3 x 148 CTAs run; slice results go to a workspace and a second kernel adds them in slice order (deterministic) and
// applies alpha / beta.
#include <algorithm>
#include "gemm.h"

#define BM 64
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
  double* ws;       // split-K workspace: slices x M x N (row stride N), or null
  int64_t sam, sak, sbk, sbn, ldc;
  int M, N, K, kchunk;
  double alpha, beta;
  const int* pred;  // optional device flag
  int pred_sense;   // 0: skip when *pred == 0 ; 1: skip when *pred != 0
};

// 8-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int nbytes = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
}

template <int BN>
__device__ __forceinline__ void load_tiles(const GemmArgs& g, double (*As)[BK + PAD], double (*Bs)[BN + PAD], int m0,
                                           int n0, int k0, int kend, bool a_kfast, bool b_nfast, int tid) {
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
    const bool ok = (gm < g.M && gk < kend);
    cp_async8(&As[m][k], ok ? g.A + (int64_t)gm * g.sam + (int64_t)gk * g.sak : g.A, ok);
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
    const bool ok = (gk < kend && gn < g.N);
    cp_async8(&Bs[k][n], ok ? g.B + (int64_t)gk * g.sbk + (int64_t)gn * g.sbn : g.B, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NT>  // n8 DMMA tiles per warp: block tile width BN = 16 * NT
__global__ void __launch_bounds__(256) k_gemm_f64(const GemmArgs g) {
  if (g.pred && ((*g.pred == 0) != (g.pred_sense != 0))) return;
  constexpr int BN = 16 * NT;
  __shared__ double As[2][BM][BK + PAD];
  __shared__ double Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
  double acc[2][NT][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const bool a_kfast = (g.sak == 1);
  const bool b_nfast = (g.sbn == 1);
  int buf = 0;
  if (kbeg < kend) load_tiles<BN>(g, As[0], Bs[0], m0, n0, kbeg, kend, a_kfast, b_nfast, tid);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // prefetch the next K tile into the other buffer while this one feeds the tensor pipe
    if (k0 + BK < kend) load_tiles<BN>(g, As[buf ^ 1], Bs[buf ^ 1], m0, n0, k0 + BK, kend, a_kfast, b_nfast, tid);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[2], b[NT];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = As[buf][wm * 16 + i * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = Bs[buf][kk + (lane & 3)][wn * (8 * NT) + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    buf ^= 1;
  }
  // ---- epilogue: C = alpha*acc + beta*C, or the raw slice into the workspace
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int m = m0 + wm * 16 + i * 8 + (lane >> 2);
      const int n = n0 + wn * (8 * NT) + j * 8 + 2 * (lane & 3);
      if (m < g.M) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (n + c < g.N) {
            if (g.ws) {
              g.ws[((size_t)blockIdx.z * g.M + m) * g.N + n + c] = acc[i][j][c];
            } else {
              double* p = g.C + (int64_t)m * g.ldc + n + c;
              const double v = g.alpha * acc[i][j][c];
              *p = (g.beta == 0.0) ? v : (v + g.beta * *p);
            }
          }
        }
      }
    }
}

__global__ void __launch_bounds__(256) k_gemm_reduce(const GemmArgs g, int slices) {
  if (g.pred && ((*g.pred == 0) != (g.pred_sense != 0))) return;
  const int64_t total = (int64_t)g.M * g.N, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    double s = 0.0;
    for (int z = 0; z < slices; ++z) s += g.ws[(size_t)z * total + i];
    const int64_t m = i / g.N, n = i % g.N;
    double* p = g.C + m * g.ldc + n;
    const double v = g.alpha * s;
    *p = (g.beta == 0.0) ? v : (v + g.beta * *p);
  }
}

static int launch(manisdp_handle* h, GemmArgs g) {
  const int BN = g.N <= 16 ? 16 : (g.N <= 32 ? 32 : 64);
  const int tm = (g.M + BM - 1) / BM, tn = (g.N + BN - 1) / BN;
  const int ksteps = (g.K + BK - 1) / BK;
  int slices = 1;
  const int target = 3 * h->num_sms;
  if (tm * tn < target && ksteps >= 16) {
    slices = std::min({(target + tm * tn - 1) / (tm * tn), ksteps / 8, 32});
    if (slices < 1) slices = 1;
  }
  g.kchunk = ((ksteps + slices - 1) / slices) * BK;
  slices = (g.K + g.kchunk - 1) / g.kchunk;
  g.ws = nullptr;
  if (slices > 1) {
    const size_t need = (size_t)slices * g.M * g.N;
    if (h->gemm_ws_cap < need) {
      // (re)allocation is synchronous; it happens only when the factor width grows, never inside a graph capture
      // because the first product of a solve runs before the graph is built
      if (h->gemm_ws) cudaFree(h->gemm_ws);
      h->gemm_ws = nullptr;
      h->gemm_ws_cap = need + need / 2;
      CUDA_TRY(h, cudaMalloc((void**)&h->gemm_ws, h->gemm_ws_cap * sizeof(double)));
    }
    g.ws = h->gemm_ws;
  }
  dim3 grid(tn, tm, slices);
  if (BN == 16)
    k_gemm_f64<1><<<grid, 256, 0, h->stream>>>(g);
  else if (BN == 32)
    k_gemm_f64<2><<<grid, 256, 0, h->stream>>>(g);
  else
    k_gemm_f64<4><<<grid, 256, 0, h->stream>>>(g);
  KERNEL_CHECK(h);
  if (slices > 1) {
    const int64_t total = (int64_t)g.M * g.N;
    const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)h->num_sms * 8));
    k_gemm_reduce<<<nb, 256, 0, h->stream>>>(g, slices);
    KERNEL_CHECK(h);
  }
  return MANISDP_OK;
}

// out(n x w, ld = ldo) = alpha * S(n x n, row stride n) * V(n x w, ld = ldv) + beta * out
int msdp_gemm_nn(manisdp_handle* h, const double* S, int n, const double* V, int ldv, int w, double* out, int ldo,
                 double alpha, double beta, const int* pred, int pred_sense) {
  GemmArgs g{S, V, out, nullptr, n, 1, ldv, 1, ldo, n, w, n, 0, alpha, beta, pred, pred_sense};
  return launch(h, g);
}

// M(n x n, row stride n) = alpha * P(n x w) * Q(n x w)'
int msdp_gemm_nt(manisdp_handle* h, const double* P, int ldp, const double* Q, int ldq, int n, int w, double* M,
                 double alpha, const int* pred, int pred_sense) {
  GemmArgs g{P, Q, M, nullptr, ldp, 1, 1, ldq, n, n, n, w, 0, alpha, 0.0, pred, pred_sense};
  return launch(h, g);
}

// G(ka x kb, row stride kb) = P(nrows x ka, ld = ldp)' * Q(nrows x kb, ld = ldq)   -- tall-skinny Gram matrices (rank step)
int msdp_gemm_tn(manisdp_handle* h, const double* P, int ldp, int ka, const double* Q, int ldq, int kb, int64_t nrows,
                 double* G) {
  GemmArgs g{P, Q, G, nullptr, 1, ldp, ldq, 1, kb, ka, kb, (int)nrows, 0, 1.0, 0.0, nullptr, 0};
  return launch(h, g);
}

// out(nrows x w, ld = ldo) = alpha * P(nrows x k, ld = ldp) * Cm(k x w, row stride ldc) + beta * out
int msdp_gemm_rows_small(manisdp_handle* h, const double* P, int ldp, int k, const double* Cm, int ldc, int w,
                         int64_t nrows, double* out, int ldo, double alpha, double beta) {
  GemmArgs g{P, Cm, out, nullptr, ldp, 1, ldc, 1, ldo, (int)nrows, w, k, 0, alpha, beta, nullptr, 0};
  return launch(h, g);
}
