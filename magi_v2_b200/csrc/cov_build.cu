// (1) Matern covariance blocks C, C', C'' for all (dataset, component) pairs.
// Replaces magi_v2.py:781-815.  One CTA per 32x32 tile of the lower triangle: 1024 Bessel
// evaluations, the mirrored tile is produced through a shared-memory transpose so that both
// tiles are written with coalesced rows (C, C'' symmetric; C' antisymmetric).
#include "bessel.cuh"
#include "common.cuh"

namespace {

constexpr int kTile = 32;

__global__ void __launch_bounds__(256)
cov_build_kernel(const double* __restrict__ I, int64_t I_stride, const double* __restrict__ phi1,
                 const double* __restrict__ phi2, MaternConsts mc, int D, int n, int flags,
                 double* __restrict__ C, double* __restrict__ Cp, double* __restrict__ Cpp, int ntile) {
  __shared__ double sC[kTile][kTile + 1], sP[kTile][kTile + 1], sQ[kTile][kTile + 1];
  // decode lower-triangular tile index
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const size_t bd = blockIdx.y;
  const int b = (int)(bd / D);
  const double* grid = I + (size_t)b * I_stride;
  const double p1 = phi1[bd], p2 = phi2[bd];
  const double h = (flags & MAGI_COV_UNIFORM_GRID) ? (grid[n - 1] - grid[0]) / (double)(n - 1) : 0.0;
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  const double diagQ = mc.nu * p1 / (p2 * p2 * (mc.nu - 1.0));
  for (int ty = ty0; ty < kTile; ty += 8) {
    const int i = ti * kTile + ty, j = tj * kTile + tx;
    double kap = 0.0, dk = 0.0, d2k = 0.0;
    if (i < n && j < n) {
      if (i == j) {
        kap = p1; dk = 0.0; d2k = -diagQ;   // limits l -> 0+  (magi_v2.py:795, :802, :815)
      } else {
        const double s = grid[i], tt = grid[j];
        const double l = (flags & MAGI_COV_UNIFORM_GRID) ? fabs((double)(i - j)) * h : fabs(s - tt);
        matern_lag(mc, p1, p2, l, kap, dk, d2k);
        dk = s > tt ? dk : -dk;  // d kappa/d s = kappa'(l) * sign(s - t)
      }
    }
    sC[ty][tx] = kap;
    sP[ty][tx] = dk;      // d kappa / d s
    sQ[ty][tx] = -d2k;    // d^2 kappa / d s d t = - kappa''(l)
    if (i < n && j < n) {
      const size_t o = bd * (size_t)n * n + (size_t)i * n + j;
      if (C) C[o] = kap;
      if (Cp) Cp[o] = dk;
      if (Cpp) Cpp[o] = -d2k;
    }
  }
  if (ti == tj) return;
  __syncthreads();
  for (int ty = ty0; ty < kTile; ty += 8) {
    // mirrored tile: element (tj*32 + ty, ti*32 + tx) = transpose of (ti*32 + tx, tj*32 + ty)
    const int i = tj * kTile + ty, j = ti * kTile + tx;
    if (i < n && j < n) {
      const size_t o = bd * (size_t)n * n + (size_t)i * n + j;
      if (C) C[o] = sC[tx][ty];
      if (Cp) Cp[o] = -sP[tx][ty];
      if (Cpp) Cpp[o] = sQ[tx][ty];
    }
  }
}

// Uniform grid: the blocks are Toeplitz, kappa(|i - j| h), so a CTA evaluates the n distinct lags once into shared
// memory (n Bessel evaluations instead of 32 x n for its 64 rows) and then only streams its rows out: the kernel
// is bound by the 24 n^2 output bytes per matrix.  Same lag arithmetic as the general kernel -> identical values.
constexpr int kRows = 64;

__global__ void __launch_bounds__(256)
cov_build_toeplitz_kernel(const double* __restrict__ I, int64_t I_stride, const double* __restrict__ phi1,
                          const double* __restrict__ phi2, MaternConsts mc, int D, int n,
                          double* __restrict__ C, double* __restrict__ Cp, double* __restrict__ Cpp) {
  extern __shared__ double tab[];  // [3][n]: kappa, d kappa/d s (s > t), d^2 kappa / d s d t
  const size_t bd = blockIdx.y;
  const int b = (int)(bd / D);
  const double* grid = I + (size_t)b * I_stride;
  const double p1 = phi1[bd], p2 = phi2[bd];
  const double h = (grid[n - 1] - grid[0]) / (double)(n - 1);
  const double diagQ = mc.nu * p1 / (p2 * p2 * (mc.nu - 1.0));
  const int r0 = blockIdx.x * kRows, r1 = min(n, r0 + kRows);
  const int lmax = max(r1 - 1, n - 1 - r0);  // largest lag this CTA's rows reach
  for (int l = threadIdx.x; l <= lmax; l += 256) {
    double kap = p1, dk = 0.0, d2k = -diagQ;
    if (l > 0) matern_lag(mc, p1, p2, fabs((double)l) * h, kap, dk, d2k);
    tab[l] = kap;
    tab[n + l] = dk;
    tab[2 * n + l] = -d2k;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = r0 + warp; i < r1; i += 8) {
    const size_t o = bd * (size_t)n * n + (size_t)i * n;
    for (int j = lane; j < n; j += 32) {
      const int l = i > j ? i - j : j - i;
      if (C) C[o + j] = tab[l];
      if (Cp) Cp[o + j] = i >= j ? tab[n + l] : -tab[n + l];
      if (Cpp) Cpp[o + j] = tab[2 * n + l];
    }
  }
}

}  // namespace

extern "C" int magi_b200_cov_build(const double* I, int64_t I_batch_stride, const double* phi1,
                                   const double* phi2, double nu, int B, int D, int n, int flags, double* C,
                                   double* Cp, double* Cpp, magi_stream_t stream) {
  if (!I) return -1;
  if (I_batch_stride != 0 && I_batch_stride < n) return -2;
  if (!phi1) return -3;
  if (!phi2) return -4;
  MaternConsts mc;
  if (matern_consts_init(nu, &mc) != 0) return -5;
  if (B <= 0) return -6;
  if (D <= 0) return -7;
  if (n <= 1) return -8;
  const int nt = (n + kTile - 1) / kTile;
  const int ntile = nt * (nt + 1) / 2;
  const size_t nmat = (size_t)B * D;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D > 65535) return MAGI_ERR_UNSUPPORTED;
  const size_t chunk = (size_t)(65535 / D) * D;  // gridDim.y limit, kept a multiple of D
  for (size_t off = 0; off < nmat; off += chunk) {
    const unsigned ny = (unsigned)((nmat - off) < chunk ? (nmat - off) : chunk);
    const size_t b0 = off / D;
    if ((flags & MAGI_COV_UNIFORM_GRID) && 3 * (size_t)n * sizeof(double) <= 48 * 1024) {
      cov_build_toeplitz_kernel<<<dim3((n + kRows - 1) / kRows, ny), 256, 3 * (size_t)n * sizeof(double), st>>>(
          I + b0 * (size_t)I_batch_stride, I_batch_stride, phi1 + off, phi2 + off, mc, D, n,
          C ? C + off * (size_t)n * n : nullptr, Cp ? Cp + off * (size_t)n * n : nullptr,
          Cpp ? Cpp + off * (size_t)n * n : nullptr);
      continue;
    }
    cov_build_kernel<<<dim3(ntile, ny), 256, 0, st>>>(
        I + b0 * (size_t)I_batch_stride, I_batch_stride, phi1 + off, phi2 + off, mc, D, n, flags,
        C ? C + off * (size_t)n * n : nullptr, Cp ? Cp + off * (size_t)n * n : nullptr,
        Cpp ? Cpp + off * (size_t)n * n : nullptr, ntile);
  }
  return magi_cuda_status(cudaGetLastError());
}
