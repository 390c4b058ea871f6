// (2) factorise + derive: from the Matern blocks (C, C', C'') of each (dataset, component) to
// the matrices the sampler needs,  C^-1,  m = C' C^-1,  K^-1 with K = C'' - m C'^T,  banded.
// Replaces magi_v2.py:818-820 + :126-128 (three SVD pseudo-inverses and two GEMMs per
// component on the CPU) + :271-274 (band_part).
//
// Design: a persistent grid; each CTA takes whole matrices and runs a blocked, GEMM-based FP64
// "mini-LAPACK" on them in global memory (operands stay L2-resident: <= 13 MB per matrix at
// n = 1281) with 56x56 shared-memory tiles (n = 161 -> 3 x 56 = 168, n = 1281 -> 23 x 56 = 1288):
//   left-looking blocked Cholesky (diagonal 56x56 blocks factorised and inverted in shared memory)
//   -> blocked triangular inverse -> A^-1 = L^-T L^-1 -> GEMMs for m and K -> same for K.
// Everything but the diagonal work is the CTA-level GEMM below, whose tile products run on the FP64 tensor
// cores (mma.sync.m8n8k4.f64 -> DMMA): one warp per 8-row strip of the tile, seven 8x8 accumulators per warp, so a
// k-step of 4 costs 8 shared-memory operand loads for 7 DMMAs (the 4x4 DFMA register tiles of round 1 needed 8 loads
// per 16 DFMAs and were bound by the shared-memory pipe: 14 % of the FP64 peak).
#include "common.cuh"

namespace {

#ifndef MAGI_FACTOR_CTAS
#define MAGI_FACTOR_CTAS 2     // resident CTAs per SM the kernel is compiled for (3: 80 registers, spills, slower)
#endif
constexpr int kNB = 56;        // block size = GEMM tile size (7 x 8)
constexpr int kKC = 16;        // GEMM k-chunk
constexpr int kFT = 32 * (kNB / 8);   // threads per CTA: one warp per 8-row strip of a tile
constexpr int kLd = kNB + 1;   // padded leading dimension of the diagonal-block arrays
constexpr int kLdG = 68;       // leading dimension of the GEMM staging tiles: 2 * 68 = 8 (mod 32) words, so the 16 lanes
                               // (g = 0..3, c = 0..3) of a half warp reading [k + c][i + g] hit 16 different bank pairs

__device__ __forceinline__ void dmma_f64(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct FactorSmem {
  double a[kKC][kLdG];
  double b[kKC][kLdG];
  double d[kNB][kLd];   // diagonal block
  double di[kNB][kLd];  // its inverse
  int info;
};

// C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C   (all kFT threads; C row-major with ldc).
// op(A)(i,k) = A[i*rsA + k*csA], op(B)(k,j) = B[k*rsB + j*csB]; one of each stride pair is 1.
// C may alias A when N <= kNB (each kNB-row strip of A is fully read before the strip is stored).
// Structure flags: the operands of the big products are triangular or the result is symmetric, and the tile loop
// skips what is known to be zero / redundant (the skipped terms are exact zeros, so the values do not change):
//   kGemmALowerT : op(A)(i,k) = 0 for k < i   (A = L^T of a lower-triangular L)      -> k starts at the tile's m0
//   kGemmALower  : op(A)(i,k) = 0 for k > i   (A lower-triangular)                   -> k ends at m0 + 64
//   kGemmBLower  : op(B)(k,j) = 0 for k < j   (B lower-triangular)                   -> k starts at n0
//   kGemmSymOut  : the result is symmetric (M == N): only tiles with n0 <= m0 are computed, the others mirrored
constexpr int kGemmALowerT = 1, kGemmALower = 2, kGemmBLower = 4, kGemmSymOut = 8;

__device__ void cta_gemm(int M, int N, int K, double alpha, const double* A, long rsA, long csA, const double* B,
                         long rsB, long csB, double beta, double* C, long ldc, FactorSmem& sm, int flags = 0) {
  // Warp w owns rows 8w .. 8w+7 of the 56x56 tile and all seven 8-column tiles; lane 4g+c holds C[8w+g][8t+2c],
  // C[8w+g][8t+2c+1] of tile t (the DMMA accumulator layout).  The next k-chunk is fetched from global memory into
  // registers while the current one is multiplied out of shared memory.
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, c = lane & 3;
  constexpr int kPer = kNB * kKC / kFT;  // 4 elements of each operand per thread and chunk
  constexpr int kT = kNB / 8;            // 7 column tiles
  // element -> (k, row/col) maps of the staging loads, fastest index along the operand's unit stride
  int ak[kPer], ai[kPer], bk[kPer], bj[kPer];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int e = tid + q * kFT;
    if (csA == 1) { ak[q] = e % kKC; ai[q] = e / kKC; } else { ai[q] = e % kNB; ak[q] = e / kNB; }
    if (csB == 1) { bj[q] = e % kNB; bk[q] = e / kNB; } else { bk[q] = e % kKC; bj[q] = e / kKC; }
  }
  for (int m0 = 0; m0 < M; m0 += kNB) {
    for (int n0 = 0; n0 < N; n0 += kNB) {
      if ((flags & kGemmSymOut) && n0 > m0) break;
      int kbeg = 0, kend = K;
      if (flags & kGemmALowerT) kbeg = max(kbeg, m0);
      if (flags & kGemmBLower) kbeg = max(kbeg, n0);
      if (flags & kGemmALower) kend = min(kend, m0 + kNB);
      double acc[kT][2];
#pragma unroll
      for (int t = 0; t < kT; ++t) acc[t][0] = acc[t][1] = 0.0;
      double pa[kPer], pb[kPer];
      auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          const int gi = m0 + ai[q], gk = k0 + ak[q];
          // (k < kend, not k < K: with 56-wide blocks a 16-deep chunk can reach past the end of a triangular
          // operand's block, into entries that are never written)
          pa[q] = (gi < M && gk < kend) ? A[gi * rsA + gk * csA] : 0.0;
          const int gj = n0 + bj[q], gk2 = k0 + bk[q];
          pb[q] = (gj < N && gk2 < kend) ? B[gk2 * rsB + gj * csB] : 0.0;
        }
      };
      fetch(kbeg);
      for (int k0 = kbeg; k0 < kend; k0 += kKC) {
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
          sm.a[ak[q]][ai[q]] = pa[q];
          sm.b[bk[q]][bj[q]] = pb[q];
        }
        __syncthreads();
        if (k0 + kKC < kend) fetch(k0 + kKC);
#pragma unroll
        for (int kk = 0; kk < kKC; kk += 4) {
          const double av = sm.a[kk + c][8 * warp + g];
#pragma unroll
          for (int t = 0; t < kT; ++t) dmma_f64(acc[t][0], acc[t][1], av, sm.b[kk + c][8 * t + g]);
        }
        __syncthreads();
      }
      const int gi = m0 + 8 * warp + g;
      if (gi < M) {
#pragma unroll
        for (int t = 0; t < kT; ++t) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int gj = n0 + 8 * t + 2 * c + q;
            if (gj >= N) continue;
            double* p = C + gi * ldc + gj;
            const double v = beta == 0.0 ? alpha * acc[t][q] : fma(alpha, acc[t][q], beta * *p);
            *p = v;
            if ((flags & kGemmSymOut) && n0 < m0) C[gj * ldc + gi] = v;
          }
        }
      }
    }
  }
  __syncthreads();
}

// In shared memory: Cholesky of the nb x nb block sm.d (lower), then sm.di = inverse of the factor.
// Non-positive pivot: records (base + column + 1) in sm.info (first failure only).
//   Factorisation: right-looking, ONE barrier per column -- the trailing update of step j divides by the pivot
//   instead of using a scaled column (d keeps L D^(1/2) until the end), so nobody waits for a square root or for the
//   scaling of a column; the columns are scaled in one pass afterwards.
//   Inverse X = L^-1 by forward substitution, all columns at once: thread (c, r) owns rows i = r (mod 4) of column c and
//   keeps their partial sums in registers; step k publishes row k of X (one barrier per step).
__device__ void smem_potrf_trtri(FactorSmem& sm, int nb, int base) {
  const int tid = threadIdx.x;
  static_assert(kFT == 4 * kNB, "trtri: four threads per column");
  for (int j = 0; j < nb; ++j) {
    __syncthreads();   // column j (and the pivot) carry every update of the steps before
    const double p = sm.d[j][j];
    if (tid == 0 && !(p > 0.0) && sm.info == 0) sm.info = base + j + 1;
    // trailing update of the lower triangle: d[i][k] -= d[i][j] d[k][j] / p,  j < k <= i < nb
    for (int i = j + 1 + (tid >> 4); i < nb; i += kFT / 16) {
      const double dij = sm.d[i][j] / p;
      for (int k = j + 1 + (tid & 15); k <= i; k += 16) sm.d[i][k] = fma(-dij, sm.d[k][j], sm.d[i][k]);
    }
  }
  __syncthreads();
  double* rs = &sm.a[0][0];   // sqrt(pivot) per column (the staging tiles are idle here)
  if (tid < nb) rs[tid] = sqrt(sm.d[tid][tid]);
  __syncthreads();
  for (int e = tid; e < nb * nb; e += kFT) {
    const int i = e / nb, j = e - i * nb;
    if (j < i) sm.d[i][j] /= rs[j];
    else if (j == i) sm.d[i][i] = rs[i];
  }
  for (int e = tid; e < kNB * kNB; e += kFT) sm.di[e / kNB][e % kNB] = 0.0;
  __syncthreads();
  // X = L^-1:  X[k][c] = (delta_kc - sum_{j<k} L[k][j] X[j][c]) / L[k][k]
  {
    const int c = tid >> 2, r = tid & 3;
    constexpr int kRows = kNB / 4;   // rows per thread
    double acc[kRows];
#pragma unroll
    for (int q = 0; q < kRows; ++q) acc[q] = 0.0;
    for (int k = 0; k < nb; ++k) {
      // publish row k of column c (its owner has the complete sum)
      if ((k & 3) == r && c < nb && c <= k) {
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < kRows; ++q)
          if (q == (k >> 2)) a = acc[q];
        sm.di[k][c] = ((c == k ? 1.0 : 0.0) - a) / rs[k];   // L[k][k] = rs[k]
      }
      __syncthreads();
      if (c < nb && c <= k) {
        const double x = sm.di[k][c];
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
          const int i = 4 * q + r;
          if (i > k && i < nb) acc[q] = fma(sm.d[i][k], x, acc[q]);
        }
      }
    }
  }
  __syncthreads();
}

// A (n x n, lower triangle used) -> L (in Lf, lower; upper part left as is) and Linv = L^-1 (block lower
// triangle of an n x n array; see below).  T: scratch kNB x n.  All in global memory.
__device__ void cta_chol_inverse(double* Lf, double* Linv, double* T, int n, FactorSmem& sm) {
  // Linv is written block row by block row: the diagonal blocks in full (zeros above the diagonal), the blocks left
  // of them by the triangular inverse below; the blocks right of the diagonal are never written NOR read (every
  // product that takes Linv as an operand carries the matching kGemm*Lower flag).
  const int tid = threadIdx.x;
  const int nblk = (n + kNB - 1) / kNB;
  for (int k = 0; k < nblk; ++k) {
    const int c0 = k * kNB, nb = min(kNB, n - c0);
    // left-looking update of block column k:  A[c0:n, c0:c0+nb] -= L[c0:n, 0:c0] L[c0:c0+nb, 0:c0]^T
    if (k > 0) cta_gemm(n - c0, nb, c0, -1.0, Lf + (size_t)c0 * n, n, 1, Lf + (size_t)c0 * n, 1, n, 1.0,
                        Lf + (size_t)c0 * n + c0, n, sm);
    for (int e = tid; e < kNB * kNB; e += kFT) {
      const int i = e / kNB, j = e % kNB;
      sm.d[i][j] = (i < nb && j < nb && j <= i) ? Lf[(size_t)(c0 + i) * n + c0 + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    smem_potrf_trtri(sm, nb, c0);
    for (int e = tid; e < nb * nb; e += kFT) {
      const int i = e / nb, j = e % nb;
      if (j <= i) Lf[(size_t)(c0 + i) * n + c0 + j] = sm.d[i][j];
      Linv[(size_t)(c0 + i) * n + c0 + j] = sm.di[i][j];
    }
    __syncthreads();
    // panel below the diagonal block:  L[r, blk] = A[r, blk] * inv(L_kk)^T   (in place, N = nb <= 64)
    const int r0 = c0 + nb;
    if (r0 < n)
      cta_gemm(n - r0, nb, nb, 1.0, Lf + (size_t)r0 * n + c0, n, 1, Linv + (size_t)c0 * n + c0, 1, n, 0.0,
               Lf + (size_t)r0 * n + c0, n, sm);
  }
  // blocked triangular inverse, block row by block row:
  //   Linv[i, 0:c0] = -Linv_ii * ( L[i, 0:c0] * Linv[0:c0, 0:c0] )
  for (int i = 1; i < nblk; ++i) {
    const int c0 = i * kNB, nb = min(kNB, n - c0);
    cta_gemm(nb, c0, c0, 1.0, Lf + (size_t)c0 * n, n, 1, Linv, n, 1, 0.0, T, n, sm, kGemmBLower);
    cta_gemm(nb, c0, nb, -1.0, Linv + (size_t)c0 * n + c0, n, 1, T, n, 1, 0.0, Linv + (size_t)c0 * n, n, sm);
  }
}

__global__ void __launch_bounds__(kFT, MAGI_FACTOR_CTAS)
factor_kernel(const double* __restrict__ C, const double* __restrict__ Cp, const double* __restrict__ Cpp, int nmat,
              int n, int band, double jitter, double* __restrict__ Cinv, double* __restrict__ m,
              double* __restrict__ Kinv, double* __restrict__ Kout, int32_t* __restrict__ info, double* ws) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smraw);
  const int tid = threadIdx.x;
  const size_t nn = (size_t)n * n;
  double* Lf = ws + (size_t)blockIdx.x * (3 * nn + (size_t)kNB * n);
  double* Linv = Lf + nn;
  double* W = Linv + nn;
  double* T = W + nn;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    const double* c = C + mat * nn;
    const double* cp = Cp + mat * nn;
    const double* cpp = Cpp + mat * nn;
    double* ci = Cinv + mat * nn;
    double* mo = m + mat * nn;
    double* ki = Kinv + mat * nn;
    if (tid == 0) sm.info = 0;
    for (size_t e = tid; e < nn; e += kFT) {
      const int i = (int)(e / n), j = (int)(e % n);
      Lf[e] = c[e] + (i == j ? jitter : 0.0);
    }
    __syncthreads();
    cta_chol_inverse(Lf, Linv, T, n, sm);
    const int infoC = sm.info;
    __syncthreads();
    if (tid == 0) sm.info = 0;
    // C^-1 = Linv^T Linv
    cta_gemm(n, n, n, 1.0, Linv, 1, n, Linv, n, 1, 0.0, ci, n, sm, kGemmALowerT | kGemmBLower | kGemmSymOut);
    // W = L^-1 C'^T.  m and K are formed from W rather than through the explicit inverse:
    //   m = C' C^-1 = W^T L^-1 ,   K = C'' - C' C^-1 C'^T = C'' - W^T W
    // -- the subtraction of a Gram matrix keeps K symmetric positive definite to rounding, where
    // C'' - (C' C^-1) C'^T loses eps * cond(C) (non-PD pivots at n = 1281 without jitter).
    cta_gemm(n, n, n, 1.0, Linv, n, 1, cp, 1, n, 0.0, W, n, sm, kGemmALower);
    cta_gemm(n, n, n, 1.0, W, 1, n, Linv, n, 1, 0.0, mo, n, sm, kGemmBLower);
    for (size_t e = tid; e < nn; e += kFT) ki[e] = cpp[e];
    __syncthreads();
    cta_gemm(n, n, n, -1.0, W, 1, n, W, n, 1, 1.0, ki, n, sm, kGemmSymOut);
    // symmetrise K (the factorisation reads the lower triangle), keep a copy if asked for
    for (size_t e = tid; e < nn; e += kFT) {
      const int i = (int)(e / n), j = (int)(e % n);
      if (j <= i) {
        const double v = 0.5 * (ki[e] + ki[(size_t)j * n + i]);
        Lf[e] = v + (i == j ? jitter : 0.0);
        if (Kout) { Kout[mat * nn + e] = v; Kout[mat * nn + (size_t)j * n + i] = v; }
      }
    }
    __syncthreads();
    cta_chol_inverse(Lf, Linv, T, n, sm);
    const int infoK = sm.info;
    cta_gemm(n, n, n, 1.0, Linv, 1, n, Linv, n, 1, 0.0, ki, n, sm, kGemmALowerT | kGemmBLower | kGemmSymOut);
    if (band >= 0) {
      for (size_t e = tid; e < nn; e += kFT) {
        const int i = (int)(e / n), j = (int)(e % n);
        if (abs(i - j) > band) { ci[e] = 0.0; mo[e] = 0.0; ki[e] = 0.0; }
      }
    }
    if (tid == 0) info[mat] = infoC ? infoC : -infoK;
    __syncthreads();
  }
}

// A -> A^-1 (full symmetric) and log det A, one matrix per CTA at a time (include/magi_b200.h: magi_b200_spd_inverse)
__global__ void __launch_bounds__(kFT, MAGI_FACTOR_CTAS)
spd_inverse_kernel(const double* __restrict__ A, int nmat, int n, double* __restrict__ Ainv,
                   double* __restrict__ logdet, int32_t* __restrict__ info, double* ws) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smraw);
  const int tid = threadIdx.x;
  const size_t nn = (size_t)n * n;
  double* Lf = ws + (size_t)blockIdx.x * (3 * nn + (size_t)kNB * n);
  double* Linv = Lf + nn;
  double* T = Linv + 2 * nn;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    if (tid == 0) sm.info = 0;
    for (size_t e = tid; e < nn; e += kFT) Lf[e] = A[mat * nn + e];
    __syncthreads();
    cta_chol_inverse(Lf, Linv, T, n, sm);
    // log det A = 2 sum_i log L_ii (fixed summation order: thread 0 adds the per-thread partial sums)
    double part = 0.0;
    for (int i = tid; i < n; i += kFT) part += log(Lf[(size_t)i * n + i]);
    double* red = &sm.a[0][0];
    red[tid] = part;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int k = 0; k < kFT; ++k) t += red[k];
      logdet[mat] = 2.0 * t;
      info[mat] = sm.info;
    }
    __syncthreads();
    cta_gemm(n, n, n, 1.0, Linv, 1, n, Linv, n, 1, 0.0, Ainv + mat * nn, n, sm, kGemmALowerT | kGemmBLower | kGemmSymOut);
  }
}

int factor_grid(int nmat) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int g = MAGI_FACTOR_CTAS * sms;   // CTAs of 7 warps and 69 KB of shared memory
  return nmat < g ? nmat : g;
}

}  // namespace

extern "C" size_t magi_b200_factor_workspace_bytes(int nmat, int n) {
  if (nmat <= 0 || n <= 0) return 0;
  // sized for the largest grid any device in the box would get (3 CTAs per SM, <= 148 SMs on B200);
  // the launch clamps its grid to what this many bytes can serve.
  const size_t per = (3 * (size_t)n * n + (size_t)kNB * n) * sizeof(double);
  const int g = nmat < 148 * MAGI_FACTOR_CTAS ? nmat : 148 * MAGI_FACTOR_CTAS;
  return per * g;
}

extern "C" int magi_b200_factor_derive(const double* C, const double* Cp, const double* Cpp, int nmat, int n,
                                       int band, double jitter, double* Cinv, double* m, double* Kinv, double* K,
                                       int32_t* info, void* workspace, size_t workspace_bytes,
                                       magi_stream_t stream) {
  if (!C) return -1;
  if (!Cp) return -2;
  if (!Cpp) return -3;
  if (nmat <= 0) return -4;
  if (n <= 1) return -5;
  if (!(jitter >= 0.0)) return -7;
  if (!Cinv) return -8;
  if (!m) return -9;
  if (!Kinv) return -10;
  if (!info) return -12;
  const size_t per = (3 * (size_t)n * n + (size_t)kNB * n) * sizeof(double);
  int grid = factor_grid(nmat);
  if (!workspace || workspace_bytes < per) return -13;
  if ((size_t)grid * per > workspace_bytes) grid = (int)(workspace_bytes / per);
  cudaError_t e = cudaFuncSetAttribute(factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FactorSmem));
  if (e != cudaSuccess) return magi_cuda_status(e);
  factor_kernel<<<grid, kFT, sizeof(FactorSmem), static_cast<cudaStream_t>(stream)>>>(
      C, Cp, Cpp, nmat, n, band, jitter, Cinv, m, Kinv, K, info, static_cast<double*>(workspace));
  return magi_cuda_status(cudaGetLastError());
}

extern "C" int magi_b200_spd_inverse(const double* A, int nmat, int n, double* Ainv, double* logdet, int32_t* info,
                                     void* workspace, size_t workspace_bytes, magi_stream_t stream) {
  if (!A) return -1;
  if (nmat <= 0) return -2;
  if (n <= 1) return -3;
  if (!Ainv) return -4;
  if (!logdet) return -5;
  if (!info) return -6;
  static_assert(kFT <= kKC * kLdG, "the reduction buffer of spd_inverse_kernel aliases the staging tile");
  const size_t per = (3 * (size_t)n * n + (size_t)kNB * n) * sizeof(double);
  int grid = factor_grid(nmat);
  if (!workspace || workspace_bytes < per) return -7;
  if ((size_t)grid * per > workspace_bytes) grid = (int)(workspace_bytes / per);
  cudaError_t e = cudaFuncSetAttribute(spd_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FactorSmem));
  if (e != cudaSuccess) return magi_cuda_status(e);
  spd_inverse_kernel<<<grid, kFT, sizeof(FactorSmem), static_cast<cudaStream_t>(stream)>>>(
      A, nmat, n, Ainv, logdet, info, static_cast<double*>(workspace));
  return magi_cuda_status(cudaGetLastError());
}
