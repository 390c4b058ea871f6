// FP64 throughput probes (include/magi_b200_probe.h): the roofs bench.py quotes for FP64-bound kernels are
// measured on the device the benchmark runs on, not assumed.
#include "common.cuh"
#include "../../include/magi_b200_probe.h"

namespace {

__global__ void probe_dfma_kernel(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void probe_dmma_kernel(double* out, int iters) {
  double c[4][2];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = 0.0;
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1])
                   : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int magi_b200_probe_fp64(int kind, int iters, int blocks, int threads, double* out, double* flops,
                                    magi_stream_t stream) {
  if (kind != 0 && kind != 1) return -1;
  if (iters <= 0) return -2;
  if (blocks <= 0) return -3;
  if (threads <= 0 || threads > 1024 || threads % 32) return -4;
  if (!out) return -5;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == 0) {
    probe_dfma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    if (flops) *flops = 2.0 * 8 * (double)iters * blocks * threads;
  } else {
    probe_dmma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    if (flops) *flops = 2.0 * 256 * 4 * (double)iters * blocks * (threads / 32);   // m8n8k4 = 256 FMA per warp
  }
  return magi_cuda_status(cudaGetLastError());
}
