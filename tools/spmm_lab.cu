// spmm_lab.cu -- standalone laboratory for the K1 inner loop (random CSR, n = 1e6, 48 nnz/row, rows of 64 doubles).
//   V0  register gathers, unroll 4 (the round-1 kernel's inner loop), one pass
//   V2  cp.async.bulk (UBLKCP) gathers into a per-warp shared-memory ring + mbarrier, metadata prefetched one row
//       ahead, B column passes with read-modify-write of the partial rows
// Prints time per product and the max difference to V0.   nvcc -arch=sm_100a -O3 -o spmm_lab spmm_lab.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <random>
#include <vector>

#define LD 64
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_v0(const int* __restrict__ rowptr, const int* __restrict__ col,
                                            const double* __restrict__ val, const double* __restrict__ U,
                                            double* __restrict__ out, long n) {
  const int lane = threadIdx.x & 31;
  const long nw = (long)gridDim.x * 8;
  for (long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += nw) {
    const int e0 = rowptr[row], e1 = rowptr[row + 1];
    double2 acc = make_double2(0, 0);
    for (int base = e0; base < e1; base += 32) {
      int c = 0; double w = 0;
      if (base + lane < e1) { c = __ldg(col + base + lane); w = __ldg(val + base + lane); }
      const int cnt = min(32, e1 - base);
      int k = 0;
      for (; k + 4 <= cnt; k += 4) {
        double2 u[4]; double ww[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int cj = __shfl_sync(0xffffffffu, c, k + s);
          ww[s] = __shfl_sync(0xffffffffu, w, k + s);
          u[s] = __ldg(reinterpret_cast<const double2*>(U + (size_t)cj * LD) + lane);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) { acc.x = fma(ww[s], u[s].x, acc.x); acc.y = fma(ww[s], u[s].y, acc.y); }
      }
      for (; k < cnt; ++k) {
        const int cj = __shfl_sync(0xffffffffu, c, k);
        const double ww = __shfl_sync(0xffffffffu, w, k);
        const double2 u = __ldg(reinterpret_cast<const double2*>(U + (size_t)cj * LD) + lane);
        acc.x = fma(ww, u.x, acc.x); acc.y = fma(ww, u.y, acc.y);
      }
    }
    reinterpret_cast<double2*>(out + (size_t)row * LD)[lane] = acc;
  }
}

// ---- bulk-async helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// V2: WARPS warps per CTA, each with a ring of 2 batches x SLOTS rows in shared memory.
template <int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32) k_v2(const int* __restrict__ bptr0, const int* __restrict__ bptr1,
                                                   const int* __restrict__ col, const double* __restrict__ val,
                                                   const double* __restrict__ U, double* __restrict__ out, long n,
                                                   int first) {
  extern __shared__ __align__(128) unsigned char smraw[];
  double* ring = reinterpret_cast<double*>(smraw);                                  // WARPS x 2 x SLOTS x LD
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + (size_t)WARPS * 2 * SLOTS * LD * 8);  // WARPS x 2
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* myring = ring + (size_t)wid * 2 * SLOTS * LD;
  uint64_t* mybar = bars + wid * 2;
  if (lane == 0) { mbar_init(&mybar[0], 1); mbar_init(&mybar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long nw = (long)gridDim.x * WARPS;
  long row = (long)blockIdx.x * WARPS + wid;
  uint32_t it = 0;  // batch counter of this warp: buffer = it & 1, parity = (it >> 1) & 1
  // metadata pipeline: (e0n, e1n) of the next row are loaded one iteration ahead
  int e0 = 0, e1 = 0;
  if (row < n) { e0 = __ldg(bptr0 + row); e1 = __ldg(bptr1 + row); }
  for (; row < n; row += nw) {
    const long rown = row + nw;
    int e0n = 0, e1n = 0;
    if (rown < n) { e0n = __ldg(bptr0 + rown); e1n = __ldg(bptr1 + rown); }
    double2 acc = first ? make_double2(0, 0) : __ldcs(reinterpret_cast<const double2*>(out + (size_t)row * LD) + lane);
    // issue batch 0 of this row, then for each batch: issue the next one, wait for the current one, accumulate
    int pos = e0;
    int c = 0; double w = 0;
    int nb = min(SLOTS, e1 - pos);
    if (nb > 0) {
      if (lane < nb) { c = __ldcs(col + pos + lane); w = __ldcs(val + pos + lane); }
      if (lane == 0) mbar_expect_tx(&mybar[it & 1], (uint32_t)nb * LD * 8);
      __syncwarp();
      if (lane < nb) bulk_g2s(myring + ((size_t)(it & 1) * SLOTS + lane) * LD, U + (size_t)c * LD, LD * 8, &mybar[it & 1]);
    }
    while (nb > 0) {
      const int posn = pos + nb;
      const int nbn = min(SLOTS, e1 - posn);
      int cn = 0; double wn = 0;
      if (nbn > 0) {
        if (lane < nbn) { cn = __ldcs(col + posn + lane); wn = __ldcs(val + posn + lane); }
        if (lane == 0) mbar_expect_tx(&mybar[(it + 1) & 1], (uint32_t)nbn * LD * 8);
        __syncwarp();
        if (lane < nbn)
          bulk_g2s(myring + ((size_t)((it + 1) & 1) * SLOTS + lane) * LD, U + (size_t)cn * LD, LD * 8, &mybar[(it + 1) & 1]);
      }
      mbar_wait(&mybar[it & 1], (it >> 1) & 1);
      const double* slot = myring + (size_t)(it & 1) * SLOTS * LD;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (s < nb) {
          const double ws = __shfl_sync(0xffffffffu, w, s);
          const double2 u = reinterpret_cast<const double2*>(slot + (size_t)s * LD)[lane];
          acc.x = fma(ws, u.x, acc.x); acc.y = fma(ws, u.y, acc.y);
        }
      }
      __syncwarp();  // everyone has read the slots before they are refilled
      ++it;
      pos = posn; nb = nbn; c = cn; w = wn;
    }
    __stcs(reinterpret_cast<double2*>(out + (size_t)row * LD) + lane, acc);
    e0 = e0n; e1 = e1n;
  }
}

__global__ void k_block_ptrs(const int* rowptr, const int* col, long n, long jrows, int B, int* bptr) {
  for (long row = (long)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (long)gridDim.x * blockDim.x) {
    int e = rowptr[row]; const int e1 = rowptr[row + 1];
    for (int b = 0; b <= B; ++b) {
      const long lim = (long)b * jrows;
      while (e < e1 && col[e] < lim) ++e;
      bptr[(size_t)b * n + row] = (b == B) ? e1 : e;
    }
  }
}

template <int WARPS, int SLOTS>
float run_v2(const int* rowptr, const int* col, const double* val, const double* U, double* out, long n, int B, int* bptr,
             int bps, int reps) {
  const size_t smem = (size_t)WARPS * 2 * SLOTS * LD * 8 + WARPS * 2 * 8;
  CK(cudaFuncSetAttribute(k_v2<WARPS, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (B > 1) {
    k_block_ptrs<<<1184, 256>>>(rowptr, col, n, (n + B - 1) / B, B, bptr);
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto once = [&]() {
    for (int p = 0; p < B; ++p) {
      const int* b0 = B > 1 ? bptr + (size_t)p * n : rowptr;
      const int* b1 = B > 1 ? bptr + (size_t)(p + 1) * n : rowptr + 1;
      k_v2<WARPS, SLOTS><<<148 * bps, WARPS * 32, smem>>>(b0, b1, col, val, U, out, n, p == 0);
    }
  };
  once();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) once();
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}


// V3: row-stationary sweep.  A CTA keeps the accumulator rows of a tile (WARPS*RPW rows x 64 doubles) in shared
// memory and walks the B column blocks; all CTAs of a wave are in the same block at about the same time, so the operand
// block (n/B rows) stays L2-resident while it is used and no partial rows go back to DRAM.
template <int WARPS, int RPW, int UNR>
__global__ void __launch_bounds__(WARPS * 32, 1) k_v3(const int* __restrict__ bptr, const int* __restrict__ col,
                                                      const double* __restrict__ val, const double* __restrict__ U,
                                                      double* __restrict__ out, long n, int B) {
  extern __shared__ __align__(16) double accs[];  // WARPS*RPW x LD
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int TR = WARPS * RPW;
  double* myacc = accs + (size_t)wid * RPW * LD;
  for (long tile = blockIdx.x; tile * TR < n; tile += gridDim.x) {
    const long rowbase = tile * TR + (long)wid * RPW;
    for (int i = lane; i < RPW * LD / 2; i += 32) reinterpret_cast<double2*>(myacc)[i] = make_double2(0, 0);
    __syncwarp();
    for (int b = 0; b < B; ++b) {
      int e0 = 0, e1 = 0;
      if (lane < RPW && rowbase + lane < n) {
        e0 = __ldg(bptr + (size_t)b * n + rowbase + lane);
        e1 = __ldg(bptr + (size_t)(b + 1) * n + rowbase + lane);
      }
      // software pipeline over the warp's rows: entries of row r+1 are loaded while row r gathers
      int cn = 0; double wn = 0;
      {
        const int f0 = __shfl_sync(0xffffffffu, e0, 0), f1 = __shfl_sync(0xffffffffu, e1, 0);
        if (f0 + lane < f1) { cn = __ldcs(col + f0 + lane); wn = __ldcs(val + f0 + lane); }
      }
      for (int r = 0; r < RPW; ++r) {
        const int r0 = __shfl_sync(0xffffffffu, e0, r), r1 = __shfl_sync(0xffffffffu, e1, r);
        int c = cn; double w = wn;
        if (r + 1 < RPW) {
          const int f0 = __shfl_sync(0xffffffffu, e0, r + 1), f1 = __shfl_sync(0xffffffffu, e1, r + 1);
          cn = 0; wn = 0;
          if (f0 + lane < f1) { cn = __ldcs(col + f0 + lane); wn = __ldcs(val + f0 + lane); }
        }
        if (r1 <= r0) continue;
        double2 acc = reinterpret_cast<double2*>(myacc + (size_t)r * LD)[lane];
        for (int base = r0; base < r1; base += 32) {
          if (base > r0) { c = 0; w = 0; if (base + lane < r1) { c = __ldcs(col + base + lane); w = __ldcs(val + base + lane); } }
          const int cnt = min(32, r1 - base);
          for (int k = 0; k < cnt; k += UNR) {
            double2 u[UNR]; double ww[UNR];
#pragma unroll
            for (int s = 0; s < UNR; ++s) {
              if (k + s < cnt) {
                const int cj = __shfl_sync(0xffffffffu, c, k + s);
                ww[s] = __shfl_sync(0xffffffffu, w, k + s);
                u[s] = __ldg(reinterpret_cast<const double2*>(U + (size_t)cj * LD) + lane);
              } else { ww[s] = 0; u[s] = make_double2(0, 0); }
            }
#pragma unroll
            for (int s = 0; s < UNR; ++s) { acc.x = fma(ww[s], u[s].x, acc.x); acc.y = fma(ww[s], u[s].y, acc.y); }
          }
        }
        reinterpret_cast<double2*>(myacc + (size_t)r * LD)[lane] = acc;
      }
    }
    __syncwarp();
    for (int r = 0; r < RPW; ++r)
      if (rowbase + r < n) __stcs(reinterpret_cast<double2*>(out + (size_t)(rowbase + r) * LD) + lane,
                                  reinterpret_cast<double2*>(myacc + (size_t)r * LD)[lane]);
    __syncwarp();
  }
}

template <int WARPS, int RPW, int UNR>
float run_v3(const int* rowptr, const int* col, const double* val, const double* U, double* out, long n, int B, int* bptr,
             int reps) {
  const size_t smem = (size_t)WARPS * RPW * LD * 8;
  CK(cudaFuncSetAttribute(k_v3<WARPS, RPW, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_block_ptrs<<<1184, 256>>>(rowptr, col, n, (n + B - 1) / B, B, bptr);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_v3<WARPS, RPW, UNR><<<148, WARPS * 32, smem>>>(bptr, col, val, U, out, n, B);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) k_v3<WARPS, RPW, UNR><<<148, WARPS * 32, smem>>>(bptr, col, val, U, out, n, B);
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000; const int deg = 48;
  std::vector<int> rp(n + 1), ci((size_t)n * deg); std::vector<double> va((size_t)n * deg);
  std::mt19937_64 rng(1);
  for (long i = 0; i < n; ++i) {
    rp[i] = (int)(i * deg);
    int* c = &ci[(size_t)i * deg];
    for (int k = 0; k < deg; ++k) c[k] = (int)(rng() % n);
    std::sort(c, c + deg);
    for (int k = 0; k < deg; ++k) va[(size_t)i * deg + k] = (double)((rng() % 7) + 1) * 0.25;
  }
  rp[n] = (int)(n * deg);
  int *drp, *dci, *bptr; double *dva, *U, *o0, *o1;
  CK(cudaMalloc(&drp, (n + 1) * 4)); CK(cudaMalloc(&dci, ci.size() * 4)); CK(cudaMalloc(&dva, va.size() * 8));
  CK(cudaMalloc(&U, n * LD * 8)); CK(cudaMalloc(&o0, n * LD * 8)); CK(cudaMalloc(&o1, n * LD * 8));
  CK(cudaMalloc(&bptr, (size_t)17 * n * 4));
  CK(cudaMemcpy(drp, rp.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dci, ci.data(), ci.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dva, va.data(), va.size() * 8, cudaMemcpyHostToDevice));
  std::vector<double> hu((size_t)n * LD);
  for (auto& x : hu) x = (double)(rng() % 1000) / 1000.0 - 0.5;
  CK(cudaMemcpy(U, hu.data(), hu.size() * 8, cudaMemcpyHostToDevice));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_v0<<<1184, 256>>>(drp, dci, dva, U, o0, n); CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k_v0<<<1184, 256>>>(drp, dci, dva, U, o0, n);
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms0; cudaEventElapsedTime(&ms0, a, b); ms0 /= 5;
  printf("{\"variant\": \"v0_regs_unroll4\", \"ms\": %.3f}\n", ms0);
  std::vector<double> h0((size_t)n * LD), h1((size_t)n * LD);
  CK(cudaMemcpy(h0.data(), o0, h0.size() * 8, cudaMemcpyDeviceToHost));
  auto check = [&](const char* name, int B, int bps, float ms) {
    CK(cudaMemcpy(h1.data(), o1, h1.size() * 8, cudaMemcpyDeviceToHost));
    double md = 0; for (size_t i = 0; i < h0.size(); i += 97) md = std::max(md, fabs(h0[i] - h1[i]));
    printf("{\"variant\": \"%s\", \"passes\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"maxdiff\": %.2e}\n", name, B, bps, ms, md);
    fflush(stdout);
  };
  if (argc > 2) {  // row-stationary sweep only
    for (int B : {1, 4, 8, 12, 16}) {
      check("v3_rs_w16_r24_u8", B, 1, run_v3<16, 24, 8>(drp, dci, dva, U, o1, n, B, bptr, 5));
      check("v3_rs_w32_r12_u8", B, 1, run_v3<32, 12, 8>(drp, dci, dva, U, o1, n, B, bptr, 5));
      check("v3_rs_w32_r12_u4", B, 1, run_v3<32, 12, 4>(drp, dci, dva, U, o1, n, B, bptr, 5));
      check("v3_rs_w24_r16_u8", B, 1, run_v3<24, 16, 8>(drp, dci, dva, U, o1, n, B, bptr, 5));
    }
    return 0;
  }
  for (int B : {1, 4, 6, 8, 12}) {
    check("v2_bulk_w8_s8", B, 3, run_v2<8, 8>(drp, dci, dva, U, o1, n, B, bptr, 3, 5));
    check("v2_bulk_w8_s8", B, 5, run_v2<8, 8>(drp, dci, dva, U, o1, n, B, bptr, 5, 5));
    check("v2_bulk_w16_s8", B, 2, run_v2<16, 8>(drp, dci, dva, U, o1, n, B, bptr, 2, 5));
    check("v2_bulk_w8_s16", B, 2, run_v2<8, 16>(drp, dci, dva, U, o1, n, B, bptr, 2, 5));
    check("v2_bulk_w4_s16", B, 6, run_v2<4, 16>(drp, dci, dva, U, o1, n, B, bptr, 6, 5));
  }
  return 0;
}
