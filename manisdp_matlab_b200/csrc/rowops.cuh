// rowops.cuh -- device helpers: row-group (sub-warp) reductions, deterministic block reduction, last-block ticket
#pragma once
#include "common.cuh"

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// lane mask of the GS-lane row group the calling lane belongs to (groups of one warp may diverge from each other)
template <int GS>
__device__ __forceinline__ unsigned group_mask() {
  if (GS == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << GS) - 1u) << (lane & ~(unsigned)(GS - 1));
}

// sum over the GS lanes of a row group (butterfly: every lane ends with the same bits)
template <int GS>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int off = GS / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, GS);
  return v;
}

// per-row manifold switch of multi-block points: the row multiplier of a Euclidean row is zero
__device__ __forceinline__ double rowsel(double v, bool oblique_row) { return oblique_row ? v : 0.0; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Deterministic block-wide sums of NQ values (fixed tree => bitwise reproducible for a fixed launch geometry).
// Result valid in thread 0.  `sm` must hold NQ * 32 doubles.
template <int NQ>
__device__ __forceinline__ void block_sum(double (&v)[NQ], double* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double s = warp_sum(v[q]);
    if (lane == 0) sm[q * 32 + w] = s;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double s = (lane < nw) ? sm[q * 32 + lane] : 0.0;
      v[q] = warp_sum(s);
    }
  }
  __syncthreads();
}

// Publish this block's partial sums and elect the last block of the grid.  In the last block (return value true
// for all of its threads) thread 0 holds the grid totals in tot[] -- summed in block-index order, independent of
// which block happened to finish last.  The ticket is reset for the next kernel.
template <int NQ>
__device__ __forceinline__ bool grid_sum_last(double (&v)[NQ], double* partials, unsigned int* ticket, double* sm,
                                              double (&tot)[NQ]) {
  __shared__ int s_last;
  block_sum<NQ>(v, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) partials[q * MSDP_MAX_BLOCKS + blockIdx.x] = v[q];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
      s += __ldcg(&partials[q * MSDP_MAX_BLOCKS + b]);
    acc[q] = s;
  }
  block_sum<NQ>(acc, sm);
#pragma unroll
  for (int q = 0; q < NQ; ++q) tot[q] = acc[q];
  if (threadIdx.x == 0) *ticket = 0u;
  return true;
}

// ticket without a reduction (kernels whose tail only needs "everyone is done")
__device__ __forceinline__ bool grid_last(unsigned int* ticket) {
  __shared__ int s_last2;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last2 = (t == gridDim.x - 1);
    if (s_last2) *ticket = 0u;
  }
  __syncthreads();
  return s_last2 != 0;
}
