// spmm_narrow_lab.cu -- laboratory for the K1 inner loop on NARROW operand rows (ld = 8 .. 64 doubles: the LOBPCG block,
// small factor widths).  Random CSR, n = 1e6, 48 nnz/row.  One row per group of GS lanes, one double2 per lane
// (ld = 2*GS), register gathers.  Knobs:
//   CM    entries fetched per lane and round (chunk = CM*GS entries): fewer dependent "indices -> gathers" rounds
//   U     gathers in flight per lane
//   MINB  __launch_bounds__ min blocks per SM (register cap -> occupancy)
//   PREF  indices of the next round are loaded before the gathers of the current one
//   PAD   remainder entries of a round are padded to the unroll width with (column 0, weight 0) instead of a scalar loop
// Rows have 1 + Poisson(48) entries (Erdos-Renyi profile + diagonal) unless a second argument is given (exactly 48).
// Prints ms per product per variant and the max difference to the first variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o spmm_narrow_lab spmm_narrow_lab.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int GS>
__device__ __forceinline__ unsigned gmask() {
  if (GS == 32) return 0xffffffffu;
  const int lane = threadIdx.x & 31;
  return ((1u << GS) - 1u) << (lane & ~(GS - 1));
}

// EPI: 0 plain store; 1 projection epilogue with its operands (Y row, U row, eG) loaded after the gathers (the product
// kernel of round 1); 2 the same with prefetch.global.L2 of those operands at the start of the row; 3 operands loaded
// into registers at the start of the row
template <int GS, int CM, int U, int MINB, bool PREF, bool PAD, int EPI>
__global__ void __launch_bounds__(256, MINB) k_narrow(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                      const double* __restrict__ val, const double* __restrict__ Ug,
                                                      double* __restrict__ out, long n, const double* __restrict__ Y,
                                                      const double* __restrict__ eG) {
  constexpr int LD = 2 * GS, CH = CM * GS;
  const unsigned mask = gmask<GS>();
  const int gl = threadIdx.x % GS;
  const long ngroups = (long)gridDim.x * (256 / GS);
  for (long row = (long)blockIdx.x * (256 / GS) + threadIdx.x / GS; row < n; row += ngroups) {
    const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
    double2 acc = make_double2(0.0, 0.0);
    double2 y3, u3;
    double eg3 = 0.0;
    if (EPI == 2) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Y + (size_t)row * LD + 2 * gl));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Ug + (size_t)row * LD + 2 * gl));
      if (gl == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(eG + row));
    } else if (EPI == 3) {
      y3 = *reinterpret_cast<const double2*>(Y + (size_t)row * LD + 2 * gl);
      u3 = *reinterpret_cast<const double2*>(Ug + (size_t)row * LD + 2 * gl);
      eg3 = eG[row];
    }
    int c[CM], cn[CM];
    double w[CM], wn[CM];
#pragma unroll
    for (int j = 0; j < CM; ++j) {
      const int e = e0 + j * GS + gl;
      c[j] = 0; w[j] = 0.0;
      if (e < e1) { c[j] = __ldg(col + e); w[j] = __ldg(val + e); }
    }
    for (int base = e0; base < e1; base += CH) {
      if (PREF) {
#pragma unroll
        for (int j = 0; j < CM; ++j) {
          const int e = base + CH + j * GS + gl;
          cn[j] = 0; wn[j] = 0.0;
          if (e < e1) { cn[j] = __ldg(col + e); wn[j] = __ldg(val + e); }
        }
      }
#pragma unroll
      for (int j = 0; j < CM; ++j) {
        int cnt = min(GS, e1 - base - j * GS);  // may be <= 0
        // PAD: lanes past the end of the row hold (column 0, weight 0), so the count is rounded up to the unroll width
        // and the one-by-one remainder loop (a full memory latency per entry) disappears
        if (PAD && cnt > 0) cnt = min(GS, (cnt + U - 1) / U * U);
        int k = 0;
        for (; k + U <= cnt; k += U) {
          double2 u[U];
          double ww[U];
#pragma unroll
          for (int s = 0; s < U; ++s) {
            const int cj = __shfl_sync(mask, c[j], k + s, GS);
            ww[s] = __shfl_sync(mask, w[j], k + s, GS);
            u[s] = __ldg(reinterpret_cast<const double2*>(Ug + (size_t)cj * LD) + gl);
          }
#pragma unroll
          for (int s = 0; s < U; ++s) { acc.x = fma(ww[s], u[s].x, acc.x); acc.y = fma(ww[s], u[s].y, acc.y); }
        }
        for (; k < cnt; ++k) {
          const int cj = __shfl_sync(mask, c[j], k, GS);
          const double ww = __shfl_sync(mask, w[j], k, GS);
          const double2 u = __ldg(reinterpret_cast<const double2*>(Ug + (size_t)cj * LD) + gl);
          acc.x = fma(ww, u.x, acc.x); acc.y = fma(ww, u.y, acc.y);
        }
      }
      if (PREF) {
#pragma unroll
        for (int j = 0; j < CM; ++j) { c[j] = cn[j]; w[j] = wn[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < CM; ++j) {
          const int e = base + CH + j * GS + gl;
          c[j] = 0; w[j] = 0.0;
          if (e < e1) { c[j] = __ldg(col + e); w[j] = __ldg(val + e); }
        }
      }
    }
    if (EPI == 0) {
      reinterpret_cast<double2*>(out + (size_t)row * LD)[gl] = acc;
    } else {
      if (EPI != 3) {
        y3 = *reinterpret_cast<const double2*>(Y + (size_t)row * LD + 2 * gl);
        u3 = *reinterpret_cast<const double2*>(Ug + (size_t)row * LD + 2 * gl);
      }
      double dot = y3.x * acc.x + y3.y * acc.y;
#pragma unroll
      for (int o = GS / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(mask, dot, o, GS);
      if (EPI != 3) eg3 = eG[row];
      double2 hv;
      hv.x = acc.x - y3.x * dot - u3.x * eg3;
      hv.y = acc.y - y3.y * dot - u3.y * eg3;
      reinterpret_cast<double2*>(out + (size_t)row * LD)[gl] = hv;
    }
  }
}

struct Ctx {
  int *rp, *ci;
  double *va, *U, *o0, *o1, *Y, *eG;
  long n, nnz;
  int sms;
  std::vector<double> h0, h1;
};

template <int GS, int CM, int U, int MINB, bool PREF, bool PAD, int EPI>
void run(Ctx& x, bool first) {
  auto kern = k_narrow<GS, CM, U, MINB, PREF, PAD, EPI>;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  const int grid = x.sms * std::max(1, occ);
  double* o = first ? x.o0 : x.o1;
  kern<<<grid, 256>>>(x.rp, x.ci, x.va, x.U, o, x.n, x.Y, x.eG);
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) kern<<<grid, 256>>>(x.rp, x.ci, x.va, x.U, o, x.n, x.Y, x.eG);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const size_t cnt = (size_t)x.n * 2 * GS;
  double md = 0.0;
  if (first) {
    CK(cudaMemcpy(x.h0.data(), x.o0, cnt * 8, cudaMemcpyDeviceToHost));
  } else {
    CK(cudaMemcpy(x.h1.data(), x.o1, cnt * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < cnt; i += 37) md = std::max(md, fabs(x.h0[i] - x.h1[i]));
  }
  printf("{\"ld\": %d, \"cm\": %d, \"u\": %d, \"minb\": %d, \"pref\": %d, \"pad\": %d, \"epi\": %d, \"regs\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, "
         "\"gather_TBps\": %.2f, \"maxdiff\": %.1e}\n",
         2 * GS, CM, U, MINB, (int)PREF, (int)PAD, EPI, fa.numRegs, occ, ms, (double)x.nnz * 2 * GS * 8 / ms / 1e9, md);
  fflush(stdout);
}

template <int GS>
void sweep(Ctx& x) {
  run<GS, 1, 4, 1, false, false, 1>(x, true);   // the round-1 product kernel: loop + late epilogue operands
  run<GS, 1, 4, 1, false, false, 0>(x, false);  // (different output: plain store; maxdiff is meaningless here)
  run<GS, 1, 4, 1, false, false, 2>(x, false);
  run<GS, 1, 4, 1, false, false, 3>(x, false);
  run<GS, 1, 4, 4, false, false, 3>(x, false);
  run<GS, 1, 4, 1, false, true, 1>(x, false);
  run<GS, 1, 4, 1, false, true, 2>(x, false);
  run<GS, 1, 4, 4, false, true, 3>(x, false);
  run<GS, 1, 4, 5, true, true, 1>(x, false);
  run<GS, 1, 4, 5, true, true, 2>(x, false);
  run<GS, 1, 4, 4, true, true, 3>(x, false);
  run<GS, 1, 8, 4, false, true, 2>(x, false);
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  const int deg = 48;
  Ctx x;
  x.n = n;
  std::vector<int> rp(n + 1), ci;
  std::vector<double> va;
  std::mt19937_64 rng(1);
  std::poisson_distribution<int> pois(deg);
  const bool exact = argc > 2;  // any second argument: exactly 48 entries per row (the first lab's setting)
  ci.reserve((size_t)n * (deg + 2));
  va.reserve((size_t)n * (deg + 2));
  std::vector<int> c;
  for (long i = 0; i < n; ++i) {
    rp[i] = (int)ci.size();
    const int len = exact ? deg : 1 + pois(rng);
    c.resize(len);
    for (int k = 0; k < len; ++k) c[k] = (int)(rng() % n);
    std::sort(c.begin(), c.end());
    for (int k = 0; k < len; ++k) {
      ci.push_back(c[k]);
      va.push_back((double)((rng() % 7) + 1) * 0.25);
    }
  }
  rp[n] = (int)ci.size();
  x.nnz = (long)ci.size();
  const size_t maxel = (size_t)n * 64;
  CK(cudaMalloc(&x.rp, (n + 1) * 4));
  CK(cudaMalloc(&x.ci, ci.size() * 4));
  CK(cudaMalloc(&x.va, va.size() * 8));
  CK(cudaMalloc(&x.U, maxel * 8));
  CK(cudaMalloc(&x.o0, maxel * 8));
  CK(cudaMalloc(&x.o1, maxel * 8));
  CK(cudaMalloc(&x.Y, maxel * 8));
  CK(cudaMalloc(&x.eG, n * 8));
  CK(cudaMemset(x.eG, 0, n * 8));
  CK(cudaMemcpy(x.rp, rp.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x.ci, ci.data(), ci.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x.va, va.data(), va.size() * 8, cudaMemcpyHostToDevice));
  std::vector<double> hu(maxel);
  for (auto& v : hu) v = (double)(rng() % 1000) / 1000.0 - 0.5;
  CK(cudaMemcpy(x.U, hu.data(), maxel * 8, cudaMemcpyHostToDevice));
  for (auto& v : hu) v = (double)(rng() % 1000) / 1000.0 - 0.5;
  CK(cudaMemcpy(x.Y, hu.data(), maxel * 8, cudaMemcpyHostToDevice));
  x.h0.resize(maxel);
  x.h1.resize(maxel);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  x.sms = prop.multiProcessorCount;
  sweep<4>(x);
  sweep<8>(x);
  sweep<16>(x);
  sweep<32>(x);
  return 0;
}
