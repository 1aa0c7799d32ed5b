// spmm_blockmajor_lab.cu -- laboratory for the open item of DESIGN.md section 4: a BLOCK-MAJOR entry format for the K1
// product on graphs without locality (n = 1e6, rows of 1 + Poisson(48) entries, operand rows of 64 doubles = 512 MB,
// 4x the L2).
//
// The operand rows are split into B column blocks of n/B rows (64 MB at B = 8: L2-resident).  The entries are stored
// block-major: all entries whose column lies in block 0 (row by row, columns ascending), then block 1, ...; each entry
// carries (col, val, row).  Pass b streams the entries of block b: a warp takes a row-aligned chunk of ~224 consecutive
// entries, i.e. hundreds of INDEPENDENT gathers from an L2-resident block (the column passes of spmm.cu had ~6 dependent
// entries per row and pass), accumulates while the row stays the same and adds the finished partial row to `out` when
// the row changes (plain read-modify-write: within a pass a row belongs to exactly one warp, passes are separate
// launches, so no atomics and a fixed summation order).  The old partial row is requested when the row starts.
//
// V0 = the round-1 single-pass register-gather loop for comparison.  Prints ms per product (all passes + the memset of
// `out`) and the max difference to V0.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o spmm_blockmajor_lab ...
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <random>
#include <vector>

#define LD 64
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_v0(const int* __restrict__ rowptr, const int* __restrict__ col,
                                            const double* __restrict__ val, const double* __restrict__ Ug,
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
          u[s] = __ldg(reinterpret_cast<const double2*>(Ug + (size_t)cj * LD) + lane);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) { acc.x = fma(ww[s], u[s].x, acc.x); acc.y = fma(ww[s], u[s].y, acc.y); }
      }
      for (; k < cnt; ++k) {
        const int cj = __shfl_sync(0xffffffffu, c, k);
        const double ww = __shfl_sync(0xffffffffu, w, k);
        const double2 u = __ldg(reinterpret_cast<const double2*>(Ug + (size_t)cj * LD) + lane);
        acc.x = fma(ww, u.x, acc.x); acc.y = fma(ww, u.y, acc.y);
      }
    }
    reinterpret_cast<double2*>(out + (size_t)row * LD)[lane] = acc;
  }
}

// one pass over the entry stream of one column block; chunk_ptr[ch] .. chunk_ptr[ch+1] = entries of chunk ch.
// Register diet: only the U gathered vectors stay live across the loads; weights are broadcast when consumed and row
// changes are a ballot mask computed once per group of 32 entries.
template <int U, int MINB, bool RMW>
__global__ void __launch_bounds__(256, MINB) k_bm(const int* __restrict__ ecol, const double* __restrict__ eval_,
                                                  const int* __restrict__ erow, const int* __restrict__ chunk_ptr,
                                                  int nchunks, const double* __restrict__ Ug, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * 8;
  for (int ch = blockIdx.x * 8 + (threadIdx.x >> 5); ch < nchunks; ch += nw) {
    const int e0 = __ldg(chunk_ptr + ch), e1 = __ldg(chunk_ptr + ch + 1);
    int cur = -1, prev_last = -1;
    double2 acc = make_double2(0.0, 0.0), old = make_double2(0.0, 0.0);
    int c = 0, r = -1;
    double w = 0.0;
    if (e0 + lane < e1) { c = __ldg(ecol + e0 + lane); w = __ldg(eval_ + e0 + lane); r = __ldg(erow + e0 + lane); }
    for (int base = e0; base < e1; base += 32) {
      int cn = 0, rn = -1;
      double wn = 0.0;
      if (base + 32 + lane < e1) {
        cn = __ldg(ecol + base + 32 + lane); wn = __ldg(eval_ + base + 32 + lane); rn = __ldg(erow + base + 32 + lane);
      }
      const int cnt = min(32, e1 - base);
      int rprev = __shfl_up_sync(0xffffffffu, r, 1);
      if (lane == 0) rprev = prev_last;
      const unsigned chg = __ballot_sync(0xffffffffu, lane < cnt && r != rprev);
      prev_last = __shfl_sync(0xffffffffu, r, cnt - 1);
      // entries past cnt carry weight 0 and column 0: the count is rounded up to the unroll width
      const int cnt_pad = min(32, (cnt + U - 1) / U * U);
      for (int k = 0; k < cnt_pad; k += U) {
        double2 u[U];
#pragma unroll
        for (int s = 0; s < U; ++s) {
          const int cj = __shfl_sync(0xffffffffu, c, k + s);
          u[s] = __ldg(reinterpret_cast<const double2*>(Ug + (size_t)cj * LD) + lane);
        }
#pragma unroll
        for (int s = 0; s < U; ++s) {
          if ((chg >> (k + s)) & 1u) {  // warp-uniform
            if (cur >= 0) {
              if (RMW) { old.x += acc.x; old.y += acc.y; } else { old = acc; }
              reinterpret_cast<double2*>(out + (size_t)cur * LD)[lane] = old;
            }
            cur = __shfl_sync(0xffffffffu, r, k + s);
            acc = make_double2(0.0, 0.0);
            if (RMW) old = reinterpret_cast<const double2*>(out + (size_t)cur * LD)[lane];
          }
          const double ws = __shfl_sync(0xffffffffu, w, k + s);
          acc.x = fma(ws, u[s].x, acc.x);
          acc.y = fma(ws, u[s].y, acc.y);
        }
      }
      c = cn; w = wn; r = rn;
    }
    if (cur >= 0) {
      if (RMW) { old.x += acc.x; old.y += acc.y; } else { old = acc; }
      reinterpret_cast<double2*>(out + (size_t)cur * LD)[lane] = old;
    }
  }
}

// Product-shaped variant of the pass for the next round (runtime row length ld <= 64, as spmm.cu needs it): the
// gathers stay UNCONDITIONAL -- lanes past ld/2 read column offset 0 of the same row and simply never store -- so the
// compiler keeps the eight loads in eight register quads back to back (check: cuobjdump -sass shows
// LDG.E.128.CONSTANT R16, R20, ... with no IMAD.MOV between load and DFMA).  The product kernel of round 1 wrote
// `u[s] = act ? ldg2(..) : 0`, which routes every load through a temporary and serialises them in ~3 groups.
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_bm_rt(const int* __restrict__ ecol, const double* __restrict__ eval_,
                                                     const int* __restrict__ erow, const int* __restrict__ chunk_ptr,
                                                     int nchunks, const double* __restrict__ Ug, double* __restrict__ part,
                                                     int ld) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < ld / 2;
  const int lo = act ? 2 * lane : 0;
  const int nw = gridDim.x * 8;
  for (int ch = blockIdx.x * 8 + (threadIdx.x >> 5); ch < nchunks; ch += nw) {
    const int e0 = __ldg(chunk_ptr + ch), e1 = __ldg(chunk_ptr + ch + 1);
    int cur = -1, prev_last = -1;
    double2 acc = make_double2(0.0, 0.0);
    int c = 0, r = -1;
    double w = 0.0;
    if (e0 + lane < e1) { c = __ldg(ecol + e0 + lane); w = __ldg(eval_ + e0 + lane); r = __ldg(erow + e0 + lane); }
    for (int base = e0; base < e1; base += 32) {
      int cn = 0, rn = -1;
      double wn = 0.0;
      if (base + 32 + lane < e1) {
        cn = __ldg(ecol + base + 32 + lane); wn = __ldg(eval_ + base + 32 + lane); rn = __ldg(erow + base + 32 + lane);
      }
      const int cnt = min(32, e1 - base);
      int rprev = __shfl_up_sync(0xffffffffu, r, 1);
      if (lane == 0) rprev = prev_last;
      const unsigned chg = __ballot_sync(0xffffffffu, lane < cnt && r != rprev);
      prev_last = __shfl_sync(0xffffffffu, r, cnt - 1);
      const int cnt_pad = min(32, (cnt + U - 1) / U * U);
      for (int k = 0; k < cnt_pad; k += U) {
        double2 u[U];
#pragma unroll
        for (int s = 0; s < U; ++s) {
          const int cj = __shfl_sync(0xffffffffu, c, k + s);
          u[s] = __ldg(reinterpret_cast<const double2*>(Ug + (size_t)cj * ld + lo));
        }
#pragma unroll
        for (int s = 0; s < U; ++s) {
          if ((chg >> (k + s)) & 1u) {
            if (cur >= 0 && act) *reinterpret_cast<double2*>(part + (size_t)cur * ld + lo) = acc;
            cur = __shfl_sync(0xffffffffu, r, k + s);
            acc = make_double2(0.0, 0.0);
          }
          const double ws = __shfl_sync(0xffffffffu, w, k + s);
          acc.x = fma(ws, u[s].x, acc.x);
          acc.y = fma(ws, u[s].y, acc.y);
        }
      }
      c = cn; w = wn; r = rn;
    }
    if (cur >= 0 && act) *reinterpret_cast<double2*>(part + (size_t)cur * ld + lo) = acc;
  }
}
// (launched by run_rt below; its summing pass reads all B partial rows of every row unconditionally from buffers that
// were zeroed once -- a (block, row) slot without entries is never written, so it stays zero -- instead of testing a mask)
template <int NB>
__global__ void __launch_bounds__(256) k_sum_all(const double* __restrict__ part, long n, int ld, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < ld / 2;
  const int lo = act ? 2 * lane : 0;
  const long nw = (long)gridDim.x * 8;
  const size_t pstride = (size_t)n * ld;
  for (long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += nw) {
    double2 v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = __ldcs(reinterpret_cast<const double2*>(part + b * pstride + (size_t)row * ld + lo));
    double2 acc = v[0];
#pragma unroll
    for (int b = 1; b < NB; ++b) { acc.x += v[b].x; acc.y += v[b].y; }
    if (act) *reinterpret_cast<double2*>(out + (size_t)row * ld + lo) = acc;
  }
}

// !RMW variant: pass b stores its partial rows into part[b] (no read, so no dependent load in the pass); this pass adds
// the partial rows a row actually has (bit b of mask[row]) -- in the product it would carry the projection epilogue
__global__ void __launch_bounds__(256) k_sum(const double* __restrict__ part, const unsigned* __restrict__ mask, int B,
                                             long n, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long nw = (long)gridDim.x * 8;
  for (long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += nw) {
    const unsigned m = __ldg(mask + row);
    double2 acc = make_double2(0.0, 0.0);
    for (int b = 0; b < B; ++b)
      if ((m >> b) & 1u) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(part + ((size_t)b * n + row) * LD) + lane);
        acc.x += v.x; acc.y += v.y;
      }
    reinterpret_cast<double2*>(out + (size_t)row * LD)[lane] = acc;
  }
}

struct BlockMajor {
  int B = 0;
  int *ecol = nullptr, *erow = nullptr, *chunk_ptr = nullptr;
  double* eval_ = nullptr;
  unsigned* mask = nullptr;  // per row: which blocks hold entries of it
  std::vector<int> chunk_off;  // per block: first chunk index (B + 1 entries)
};

static BlockMajor build(const std::vector<int>& rp, const std::vector<int>& ci, const std::vector<double>& va, long n,
                        int B, int chunk_target) {
  BlockMajor bm;
  bm.B = B;
  const long nnz = (long)ci.size();
  const long jrows = (n + B - 1) / B;
  std::vector<long> cntb(B + 1, 0);
  for (long e = 0; e < nnz; ++e) cntb[ci[e] / jrows + 1]++;
  for (int b = 0; b < B; ++b) cntb[b + 1] += cntb[b];
  std::vector<int> ecol(nnz), erow(nnz);
  std::vector<double> ev(nnz);
  std::vector<long> pos(cntb.begin(), cntb.end() - 1);
  std::vector<unsigned> hmask(n, 0u);
  for (long i = 0; i < n; ++i)
    for (int e = rp[i]; e < rp[i + 1]; ++e) {
      const long q = pos[ci[e] / jrows]++;
      ecol[q] = ci[e]; erow[q] = (int)i; ev[q] = va[e];
      hmask[i] |= 1u << (ci[e] / jrows);
    }
  // row-aligned chunks of about chunk_target entries inside each block
  std::vector<int> cptr;
  bm.chunk_off.assign(B + 1, 0);
  for (int b = 0; b < B; ++b) {
    bm.chunk_off[b] = (int)cptr.size();
    long e = cntb[b];
    const long end = cntb[b + 1];
    while (e < end) {
      cptr.push_back((int)e);
      long f = std::min(end, e + chunk_target);
      while (f < end && erow[f] == erow[f - 1]) ++f;  // finish the row
      e = f;
    }
  }
  bm.chunk_off[B] = (int)cptr.size();
  cptr.push_back((int)nnz);
  // chunk_ptr of the last chunk of block b ends at the first chunk of block b+1 = cntb[b+1]: consistent by construction
  CK(cudaMalloc(&bm.ecol, nnz * 4)); CK(cudaMalloc(&bm.erow, nnz * 4)); CK(cudaMalloc(&bm.eval_, nnz * 8));
  CK(cudaMalloc(&bm.chunk_ptr, cptr.size() * 4));
  CK(cudaMemcpy(bm.ecol, ecol.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(bm.erow, erow.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(bm.eval_, ev.data(), nnz * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(bm.chunk_ptr, cptr.data(), cptr.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&bm.mask, n * 4));
  CK(cudaMemcpy(bm.mask, hmask.data(), n * 4, cudaMemcpyHostToDevice));
  return bm;
}

static void release(BlockMajor& bm) {
  cudaFree(bm.ecol); cudaFree(bm.erow); cudaFree(bm.eval_); cudaFree(bm.chunk_ptr); cudaFree(bm.mask);
}

// product-shaped variant (runtime ld = 64 here, B = 4): zero the partial buffers once, then passes + k_sum_all
template <int U, int MINB>
static float run_rt(const BlockMajor& bm, const double* Ug, double* out, double* part, long n, int sms, int reps,
                    int* regs, int* occ) {
  auto kern = k_bm_rt<U, MINB>;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  *regs = fa.numRegs;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, 256, 0);
  const int grid = sms * std::max(1, *occ);
  CK(cudaMemset(part, 0, (size_t)bm.B * n * LD * 8));
  auto once = [&]() {
    for (int b = 0; b < bm.B; ++b) {
      const int nch = bm.chunk_off[b + 1] - bm.chunk_off[b];
      kern<<<grid, 256>>>(bm.ecol, bm.eval_, bm.erow, bm.chunk_ptr + bm.chunk_off[b], nch, Ug,
                          part + (size_t)b * n * LD, LD);
    }
    k_sum_all<4><<<sms * 8, 256>>>(part, n, LD, out);
  };
  once();
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) once();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

template <int U, int MINB, bool RMW>
static float run_bm(const BlockMajor& bm, const double* Ug, double* out, double* part, long n, int sms, int reps,
                    int* regs, int* occ) {
  auto kern = k_bm<U, MINB, RMW>;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  *regs = fa.numRegs;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, 256, 0);
  const int grid = sms * std::max(1, *occ);
  auto once = [&]() {
    if (RMW) cudaMemsetAsync(out, 0, (size_t)n * LD * 8);
    for (int b = 0; b < bm.B; ++b) {
      const int nch = bm.chunk_off[b + 1] - bm.chunk_off[b];
      kern<<<grid, 256>>>(bm.ecol, bm.eval_, bm.erow, bm.chunk_ptr + bm.chunk_off[b], nch, Ug,
                          RMW ? out : part + (size_t)b * n * LD);
    }
    if (!RMW) k_sum<<<sms * 8, 256>>>(part, bm.mask, bm.B, n, out);
  };
  once();
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) once();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  const int deg = 48;
  std::vector<int> rp(n + 1), ci;
  std::vector<double> va;
  std::mt19937_64 rng(1);
  std::poisson_distribution<int> pois(deg);
  ci.reserve((size_t)n * (deg + 2));
  va.reserve((size_t)n * (deg + 2));
  std::vector<int> c;
  for (long i = 0; i < n; ++i) {
    rp[i] = (int)ci.size();
    const int len = 1 + pois(rng);
    c.resize(len);
    for (int k = 0; k < len; ++k) c[k] = (int)(rng() % n);
    std::sort(c.begin(), c.end());
    for (int k = 0; k < len; ++k) { ci.push_back(c[k]); va.push_back((double)((rng() % 7) + 1) * 0.25); }
  }
  rp[n] = (int)ci.size();
  const long nnz = (long)ci.size();
  int *drp, *dci;
  double *dva, *Ug, *o0, *o1;
  CK(cudaMalloc(&drp, (n + 1) * 4)); CK(cudaMalloc(&dci, nnz * 4)); CK(cudaMalloc(&dva, nnz * 8));
  CK(cudaMalloc(&Ug, n * LD * 8)); CK(cudaMalloc(&o0, n * LD * 8)); CK(cudaMalloc(&o1, n * LD * 8));
  CK(cudaMemcpy(drp, rp.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dci, ci.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dva, va.data(), nnz * 8, cudaMemcpyHostToDevice));
  std::vector<double> hu((size_t)n * LD);
  for (auto& x : hu) x = (double)(rng() % 1000) / 1000.0 - 0.5;
  CK(cudaMemcpy(Ug, hu.data(), hu.size() * 8, cudaMemcpyHostToDevice));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_v0<<<sms * 4, 256>>>(drp, dci, dva, Ug, o0, n);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k_v0<<<sms * 4, 256>>>(drp, dci, dva, Ug, o0, n);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms0;
  cudaEventElapsedTime(&ms0, a, b);
  printf("{\"variant\": \"v0_single_pass\", \"nnz\": %ld, \"ms\": %.3f}\n", nnz, ms0 / 5);
  fflush(stdout);
  std::vector<double> h0((size_t)n * LD), h1((size_t)n * LD);
  CK(cudaMemcpy(h0.data(), o0, h0.size() * 8, cudaMemcpyDeviceToHost));
  double* part = nullptr;
  CK(cudaMalloc(&part, (size_t)8 * n * LD * 8));  // partial rows of up to 8 blocks
  auto report = [&](int B, int chunk, int U, int minb, int regs, int occ, float ms, int rmw = 1) {
    CK(cudaMemcpy(h1.data(), o1, h1.size() * 8, cudaMemcpyDeviceToHost));
    double md = 0;
    for (size_t i = 0; i < h0.size(); i += 61) md = std::max(md, fabs(h0[i] - h1[i]));
    printf("{\"variant\": \"block_major\", \"B\": %d, \"chunk\": %d, \"u\": %d, \"minb\": %d, \"regs\": %d, "
           "\"blocks_per_sm\": %d, \"rmw\": %d, \"ms\": %.3f, \"maxdiff\": %.2e}\n", B, chunk, U, minb, regs, occ, rmw, ms, md);
    fflush(stdout);
  };
  {  // product-shaped variant, B = 4 (reported with rmw = 2)
    BlockMajor bm = build(rp, ci, va, n, 4, 224);
    int regs, occ;
    float ms = run_rt<8, 3>(bm, Ug, o1, part, n, sms, 5, &regs, &occ);
    report(4, 224, 8, 3, regs, occ, ms, 2);
    release(bm);
  }
  const int Bs[] = {2, 3, 4, 6, 8};
  for (int B : Bs) {
    BlockMajor bm = build(rp, ci, va, n, B, 224);
    int regs, occ;
    float ms;
    ms = run_bm<4, 4, true>(bm, Ug, o1, part, n, sms, 5, &regs, &occ); report(B, 224, 4, 4, regs, occ, ms, 1);
    ms = run_bm<4, 4, false>(bm, Ug, o1, part, n, sms, 5, &regs, &occ); report(B, 224, 4, 4, regs, occ, ms, 0);
    ms = run_bm<4, 5, false>(bm, Ug, o1, part, n, sms, 5, &regs, &occ); report(B, 224, 4, 5, regs, occ, ms, 0);
    ms = run_bm<4, 6, false>(bm, Ug, o1, part, n, sms, 5, &regs, &occ); report(B, 224, 4, 6, regs, occ, ms, 0);
    ms = run_bm<8, 4, false>(bm, Ug, o1, part, n, sms, 5, &regs, &occ); report(B, 224, 8, 4, regs, occ, ms, 0);
    ms = run_bm<8, 3, false>(bm, Ug, o1, part, n, sms, 5, &regs, &occ); report(B, 224, 8, 3, regs, occ, ms, 0);
    release(bm);
  }
  return 0;
}
