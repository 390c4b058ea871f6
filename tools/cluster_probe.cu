// What the communication skeleton of a cluster-resident sampler would cost ("residency"): 8 CTAs of a
// thread-block cluster hold one dataset's matrices between them; every matrix pass then ends with an all-gather of the
// pass's result vector (8 chains x 168 grid points = 1344 doubles, each CTA producing 1/8 of it) through distributed
// shared memory and a cluster barrier.  This probe times that step alone:
//   (a) cluster.sync only, (b) all-gather by remote stores (each CTA writes its 168 doubles into the 7 peers) + sync,
//   (c) all-gather by remote loads (each CTA reads the 7 peers' slices) + 2 syncs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cluster_probe tools/cluster_probe.cu && tools/cluster_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int kCl = 8, kVec = 1344, kSlice = kVec / kCl, kThreads = 256, kIters = 2000;

template <int MODE>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads) probe(double* out, long long* cycles) {
  __shared__ double vec[2][kVec];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  for (int i = threadIdx.x; i < 2 * kVec; i += kThreads) (&vec[0][0])[i] = rank + 0.001 * i;
  cl.sync();
  const long long t0 = clock64();
  double acc = 0.0;
  for (int it = 0; it < kIters; ++it) {
    double* cur = vec[it & 1];
    double* nxt = vec[(it + 1) & 1];
    if (MODE == 1) {   // push my slice of the next vector to everybody
      for (int i = threadIdx.x; i < kSlice * kCl; i += kThreads) {
        const int peer = i / kSlice, e = i - peer * kSlice;
        double* remote = cl.map_shared_rank(nxt, peer);
        remote[rank * kSlice + e] = cur[rank * kSlice + e] * 1.0000001 + it;
      }
    } else if (MODE == 2) {   // publish my slice locally, then pull the others
      for (int e = threadIdx.x; e < kSlice; e += kThreads) nxt[rank * kSlice + e] = cur[rank * kSlice + e] * 1.0000001 + it;
      cl.sync();
      for (int i = threadIdx.x; i < kSlice * kCl; i += kThreads) {
        const int peer = i / kSlice, e = i - peer * kSlice;
        if (peer != (int)rank) nxt[peer * kSlice + e] = cl.map_shared_rank(nxt, peer)[peer * kSlice + e];
      }
    }
    cl.sync();
    acc += nxt[(threadIdx.x * 5) % kVec];
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (t1 - t0) / kIters;
  out[blockIdx.x * kThreads + threadIdx.x] = acc;
}

int main() {
  double* out;
  long long* cyc;
  const int grid = 144;   // 18 clusters of 8
  cudaMalloc(&out, grid * kThreads * sizeof(double));
  cudaMalloc(&cyc, grid * sizeof(long long));
  long long h[144];
  const char* names[3] = {"cluster.sync only", "all-gather by remote stores + sync", "all-gather by remote loads + 2 syncs"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) probe<0><<<grid, kThreads>>>(out, cyc);
      if (mode == 1) probe<1><<<grid, kThreads>>>(out, cyc);
      if (mode == 2) probe<2><<<grid, kThreads>>>(out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mn = h[0], mx = h[0], sum = 0;
    for (int i = 0; i < grid; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    printf("%-40s cycles per step: min %lld mean %lld max %lld  -> x 12 steps per evaluation = %lld cycles\n", names[mode],
           mn, sum / grid, mx, 12 * (sum / grid));
  }
  return 0;
}
