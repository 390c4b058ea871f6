// Microbenchmark: FP64 DFMA (vector pipe) and DMMA (mma.sync m8n8k4 f64) throughput on this GPU.
// The roofline for FP64-bound kernels must be measured (MEASURED_PEAKS.json has no FP64 entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
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
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[4][2];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// streaming read bandwidth with 16-byte loads (what the mat-vec kernels do)
__global__ void read_kernel(const double2* __restrict__ in, size_t n, double* out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = __ldg(in + i);
    s += v.x + v.y;
  }
  if (s == 123.456) out[0] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 16 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  for (int threads : {256, 512, 1024}) {
    const int blocks = sms * (2048 / threads);
    const int iters = 20000;
    dfma_kernel<<<blocks, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA threads=%d blocks=%d: %.2f TFLOP/s\n", threads, blocks,
           2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    dmma_kernel<<<blocks, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 = 8*8*4 FMA per warp
    printf("DMMA threads=%d blocks=%d: %.2f TFLOP/s\n", threads, blocks,
           2.0 * 256 * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
  }
  const size_t n = (size_t)1 << 29;  // 8 GiB of double2
  double2* in;
  cudaMalloc(&in, n * sizeof(double2));
  cudaMemset(in, 0, n * sizeof(double2));
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    read_kernel<<<sms * 8, 512>>>(in, n, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("read-only stream 8 GiB: %.1f GB/s\n", n * 16.0 / (ms * 1e-3) / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
