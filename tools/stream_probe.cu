// Microbenchmark behind the design of the posterior kernels' matrix streams (not part of the product):
// how many bytes must one CTA per SM keep in flight, with 128-bit global->register loads and a large
// shared-memory carve-out, to saturate HBM on B200?  Each warp streams its own contiguous run of 512-byte
// tiles with a rolling pipeline of K outstanding loads per lane, as mma_task() does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_probe tools/stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int K>
__global__ void __launch_bounds__(1024, 1)
probe(const double2* __restrict__ in, size_t tiles_per_warp, int run_tiles, double* out) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // warp's region: contiguous tiles; runs of `run_tiles` tiles are contiguous, consecutive runs are far apart
  const size_t gw = (size_t)blockIdx.x * nw + warp;
  const double2* p = in + gw * tiles_per_warp * 32 + lane;
  double2 a[K];
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < K; ++u) a[u] = __ldg(p + (size_t)u * 32);
  for (size_t t = 0; t + K < tiles_per_warp; t += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      s = fma(a[u].x, a[u].y, s);
      a[u] = __ldg(p + (t + u + K) * 32);
    }
  }
#pragma unroll
  for (int u = 0; u < K; ++u) s += a[u].x;
  if (s == 123.456) out[0] = s + smem[run_tiles & 7];
}

template <int K>
double run(const double2* in, size_t total_tiles, int warps, int smem_kb, double* out) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t tpw = total_tiles / ((size_t)sms * warps) / K * K;
  cudaFuncSetAttribute(probe<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<K><<<sms, warps * 32, smem_kb * 1024>>>(in, tpw, 21, out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<K><<<sms, warps * 32, smem_kb * 1024>>>(in, tpw, 21, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)sms * warps * tpw * 512.0 / (ms * 1e-3) / 1e9;
}

int main() {
  const size_t bytes = (size_t)8 << 30;
  double2* in;
  double* out;
  cudaMalloc(&in, bytes);
  cudaMalloc(&out, 1024);
  cudaMemset(in, 0, bytes);
  const size_t tiles = bytes / 512;
  printf("%6s %6s %8s %10s %10s\n", "warps", "K", "smemKB", "inflightKB", "GB/s");
  for (int smem_kb : {124, 200}) {
    for (int warps : {7, 11, 16, 21, 32}) {
#define RUN(K) printf("%6d %6d %8d %10.1f %10.1f\n", warps, K, smem_kb, warps * K * 0.5, run<K>(in, tiles, warps, smem_kb, out));
      RUN(4) RUN(7) RUN(10) RUN(14) RUN(20) RUN(28)
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
