// gather_bench.cu -- how fast can B200 gather random 512-byte rows (the access pattern of K1 at p = 64)?
// table of R rows x 64 doubles; each warp sums `deg` random rows; vary table size (L2-resident vs not) and unroll.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
template <int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ T, const int* __restrict__ idx, int deg,
                                                long nrows_out, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long nw = (long)gridDim.x * (blockDim.x >> 5);
  for (long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrows_out; row += nw) {
    const int* ip = idx + row * deg;
    double2 acc = make_double2(0, 0);
    for (int b = 0; b < deg; b += 32) {
      int c = (b + lane < deg) ? __ldg(ip + b + lane) : 0;
      int cnt = min(32, deg - b);
      int k = 0;
      for (; k + UNROLL <= cnt; k += UNROLL) {
        double2 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          int j = __shfl_sync(0xffffffffu, c, k + u);
          v[u] = __ldg(reinterpret_cast<const double2*>(T + (size_t)j * 64) + lane);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
      }
      for (; k < cnt; ++k) {
        int j = __shfl_sync(0xffffffffu, c, k);
        double2 v = __ldg(reinterpret_cast<const double2*>(T + (size_t)j * 64) + lane);
        acc.x += v.x; acc.y += v.y;
      }
    }
    reinterpret_cast<double2*>(out + row * 64)[lane] = acc;
  }
}
template <int U> float run(const double* T, const int* idx, int deg, long nout, double* out, int bps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = 148 * bps;
  k_gather<U><<<grid, 256>>>(T, idx, deg, nout, out);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k_gather<U><<<grid, 256>>>(T, idx, deg, nout, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
  const int deg = 48; const long nout = 1000000;
  std::vector<int> h((size_t)nout * deg);
  double* out; cudaMalloc(&out, nout * 64 * 8);
  int* idx; cudaMalloc(&idx, h.size() * 4);
  long sizes[] = {125000, 250000, 500000, 1000000};  // 64 MB, 128 MB, 256 MB, 512 MB tables
  for (long R : sizes) {
    srand(1); for (auto& x : h) x = (int)(((long)rand() * 32768 + rand()) % R);
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    double* T; cudaMalloc(&T, R * 64 * 8); cudaMemset(T, 0, R * 64 * 8);
    for (int bps : {4, 8}) {
      float t4 = run<4>(T, idx, deg, nout, out, bps), t8 = run<8>(T, idx, deg, nout, out, bps), t16 = run<16>(T, idx, deg, nout, out, bps);
      double gb = (double)nout * deg * 512 / 1e9;
      printf("{\"table_MB\": %ld, \"blocks_per_sm\": %d, \"ms_u4\": %.3f, \"ms_u8\": %.3f, \"ms_u16\": %.3f, \"TBps_u4\": %.2f, \"TBps_u8\": %.2f, \"TBps_u16\": %.2f}\n",
             R * 512 / 1000000, bps, t4, t8, t16, gb / t4, gb / t8, gb / t16);
    }
    cudaFree(T);
  }
  return 0;
}
