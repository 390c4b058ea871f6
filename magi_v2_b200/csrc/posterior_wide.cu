// Log-posterior + gradient for few datasets: the rows of every component's matrices spread over the grid
// (include/magi_b200_wide.h).  Replaces magi_v2.py:308-348 + TF autodiff exactly like posterior_core.cuh does;
// formulas as SURVEY.md A.2 / A.3 with the packed symmetric forms S_C = sym(C^-1), S_K = sym(K^-1):
//   u = S_C x_c, v = m x_c, r = f(X, theta) - v, q = S_K r, t1 = x_c.u, t2 = r.q
//   dX_d = 2 u_d + sum_d' (d f_d'/d x_d) 2 q_d' - 2 (m^T q)_d          (then / beta, + 2 e / sigma^2, * -beta_temp / 2)
// One CTA = `rbpc` 8-row blocks of one (dataset, chain group of 8, component); its 8 warps split the tile columns of a
// block row, each warp contracting 8x8 matrix tiles with the chain group's vectors on the FP64 tensor cores
// (mma.sync.m8n8k4: A = half a tile, coalesced 256 B per warp load, also for the transposed pass; B = 4 x 8 slice of
// the vector array in shared memory, conflict-free), then a cross-warp reduction.  HBM-bound: every matrix byte is
// read once per pass by exactly one warp.
#include <stdlib.h>

#include "common.cuh"
#include "ode_models.cuh"
#include "../../include/magi_b200_wide.h"
// A user-supplied ODE system (the reference's f_vec callable, magi_v2.py:32-33, :335): magi_v2_b200/tracing.py
// traces the Python callable symbolically, writes `struct UserModel` (f and its vector-Jacobian products, as the
// structs of ode_models.cuh) and compiles THIS file alone into a library of its own with
//   -DMAGI_USER_MODEL_HEADER='"<path>/user_model.cuh"'
#ifdef MAGI_USER_MODEL_HEADER
#include MAGI_USER_MODEL_HEADER
#endif

namespace {

constexpr int kCh = 8;   // chains per group (DMMA n = 8); the warps per CTA are a template parameter (KW: 8, 16 or 24)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct Geo {
  int np, nblk, kb, rbpc, nchunk, G;  // padded n, 8-row blocks, band in blocks (<0: dense), block rows per CTA, CTAs per column of the grid, chain groups
  int NG, GS;                         // chain groups per CTA (1 or 2: they share every matrix tile the CTA loads), CTAs per (chunk, d, b) = ceil(G / NG)
  int KW;                             // warps per CTA of the three passes
  size_t per_bg;                      // workspace doubles per (dataset, chain group)
};

__host__ __device__ inline size_t off_R(const Geo&, int) { return 0; }
__host__ __device__ inline size_t off_Q(const Geo& g, int D) { return (size_t)D * g.np * kCh; }
__host__ __device__ inline size_t off_t1(const Geo& g, int D) { return 2 * (size_t)D * g.np * kCh; }
__host__ __device__ inline size_t off_t2(const Geo& g, int D) { return off_t1(g, D) + (size_t)D * g.nblk * kCh; }
__host__ __device__ inline size_t off_sse(const Geo& g, int D) { return off_t2(g, D) + (size_t)D * g.nblk * kCh; }
__host__ __device__ inline size_t off_th(const Geo& g, int D) { return off_sse(g, D) + (size_t)D * g.nblk * kCh; }

Geo make_geo(const magi_problem_t* pb, int sms) {
  Geo g;
  g.np = magi_pad8(pb->n);
  g.nblk = g.np / 8;
  g.kb = pb->band < 0 ? -1 : (pb->band + 7) >> 3;
  g.G = (pb->R + kCh - 1) / kCh;
  // Warps per CTA.  The passes are bound by the latency of their tile loads (mostly L2 hits: the chain groups of a
  // dataset read the same tiles at the same time) and a CTA's shared memory is one vector array whatever its size, so
  // large grids take more warps per CTA: measured at n = 1281 (Lorenz-96, 2 x 64 chains) 1.90 / 1.33 / 1.18 / 1.34 ms
  // per sweep with 8 / 16 / 24 / 32 warps; at n = 321 1.56 / 1.54 / 1.70; at n = 161 with 20 datasets 60 / 79 / 109 us.
  g.KW = g.np >= 1024 ? 24 : (g.np >= 320 ? 16 : 8);
  if (const char* f = getenv("MAGI_WIDE_KW")) {   // experiment knob
    const int v = atoi(f);
    if (v == 8 || v == 16 || v == 24) g.KW = v;
  }
  // Chain groups per CTA.  2 = every matrix tile loaded once per 16 chains instead of once per 8 -- measured SLOWER at
  // the shape it was meant for (Lorenz-96, n = 1281, R = 64: 2.24 ms against 1.59 ms per evaluation sweep): two vector
  // arrays are 165 KB of shared memory, i.e. one CTA of 8 warps per SM instead of two, and the register-streamed tile
  // loads lose their latency cover.  Kept behind MAGI_WIDE_NG=2 for experiments; tests cover both.
  g.NG = 1;
  if (const char* f = getenv("MAGI_WIDE_NG")) {
    if (f[0] == '2' && g.KW == 8 && g.G >= 2 &&
        (2 * (size_t)g.np * kCh + 8 * 4 * 32 + 2 * kCh * pb->P) * sizeof(double) <= 200 * 1024)
      g.NG = 2;
  }
  g.GS = (g.G + g.NG - 1) / g.NG;
  const long ctas1 = (long)g.nblk * pb->D * pb->B * g.GS;      // with one block row per CTA
  long r = ctas1 / (4L * sms);
  g.rbpc = (int)(r < 1 ? 1 : (r > g.KW ? g.KW : r));
  g.nchunk = (g.nblk + g.rbpc - 1) / g.rbpc;
  g.per_bg = off_th(g, pb->D) + (size_t)g.nblk * pb->P * kCh;
  return g;
}

__device__ __forceinline__ void jrange(const Geo& g, int I, int& lo, int& hi) {
  lo = 0;
  hi = g.nblk - 1;
  if (g.kb >= 0) {
    lo = max(lo, I - g.kb);
    hi = min(hi, I + g.kb);
  }
}

// cross-warp sum of NV values per lane; result valid in warp 0
template <int NV, int KW>
__device__ __forceinline__ void cta_reduce(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // red free (previous block row consumed)
#pragma unroll
  for (int q = 0; q < NV; ++q) red[(warp * NV + q) * 32 + lane] = v[q];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < KW; ++w) s += red[(w * NV + q) * 32 + lane];
      v[q] = s;
    }
  }
}

// sum over the 8 lanes that share (lane & 3), i.e. over the rows of a block row; result in lanes 0..3
__device__ __forceinline__ double rows_sum(double v) {
  v += magi_shfl_xor(v, 4);
  v += magi_shfl_xor(v, 8);
  v += magi_shfl_xor(v, 16);
  return v;
}

struct Args {
  magi_problem_t pb;
  Geo g;
  const double* X;
  const double* sig_pre;
  const double* th_pre;
  const double* beta_temp;
  double* lp;
  double* gX;
  double* gsig;
  double* gth;
  double* ws;
};

// Two work splits, same arithmetic.  ROWW = false (few block rows per CTA: small n / very few datasets): the CTA's 8
// warps split the tile columns of one block row and their partial products are summed through shared memory.
// ROWW = true (8 block rows per CTA: enough rows to fill the grid): each warp owns a block row and streams all of its
// tiles -- no barrier and no reduction inside the loop.
#define MAGI_WIDE_ROWS(I)                                                   \
  const int I0 = chunk * g.rbpc, I1 = min(g.nblk, I0 + g.rbpc);             \
  for (int I = ROWW ? I0 + warp : I0; I < I1; I += ROWW ? KW : 1)

// ---- pass 1: u = S_C x_c, v = m x_c, r = f - v, t1 ------------------------------------------------------------
template <class M, bool ROWW, int NG, int KW>
__global__ void __launch_bounds__(32 * KW) wide_pass1(Args a) {
  extern __shared__ double sm[];
  constexpr int D = M::D, P = M::P, kT = 32 * KW;
  const Geo& g = a.g;
  const size_t vsz = (size_t)g.np * kCh;
  double* vs = sm;                          // [NG][np][8]
  double* red = vs + NG * vsz;              // [KW][4][32]
  double* ths = red + KW * 4 * 32;          // [NG][8][P]
  // the chain groups of a dataset are adjacent in launch order: they read the same matrix rows at the same time (L2 hits)
  const int grp0 = (blockIdx.x % g.GS) * NG, chunk = blockIdx.x / g.GS, d = blockIdx.y, b = blockIdx.z;
  const int n = a.pb.n, R = a.pb.R, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double mu_d = a.pb.mu[b * D + d];
#pragma unroll
  for (int g2 = 0; g2 < NG; ++g2) {
    for (int e = tid; e < g.np * kCh; e += kT) {
      const int j = e >> 3, ch = e & 7, r = (grp0 + g2) * kCh + ch;
      vs[g2 * vsz + e] = (j < n && r < R) ? a.X[(((size_t)b * R + r) * n + j) * D + d] - mu_d : 0.0;
    }
    for (int e = tid; e < kCh * P; e += kT) {
      const int ch = e / P, k = e % P, r = (grp0 + g2) * kCh + ch;
      ths[g2 * kCh * P + e] = r < R ? magi_softplus(a.th_pre[((size_t)b * R + r) * P + k]) : 1.0;
    }
  }
  const double* matC = static_cast<const double*>(a.pb.packed) + ((size_t)(b * D + d) * 3 + 0) * g.np * g.np;
  const double* matM = matC + (size_t)g.np * g.np;
  __syncthreads();
  MAGI_WIDE_ROWS(I) {
    int lo, hi;
    jrange(g, I, lo, hi);
    double c[NG][4];
#pragma unroll
    for (int g2 = 0; g2 < NG; ++g2) c[g2][0] = c[g2][1] = c[g2][2] = c[g2][3] = 0.0;
#pragma unroll 4
    for (int J = ROWW ? lo : lo + warp; J <= hi; J += ROWW ? 1 : KW) {
      const size_t t = ((size_t)I * g.nblk + J) * 64;
      const int e0 = magi_tile_pos(lane >> 2, lane & 3), e1 = magi_tile_pos(lane >> 2, 4 + (lane & 3));
      const double a0 = matC[t + e0], a1 = matC[t + e1], m0 = matM[t + e0], m1 = matM[t + e1];
#pragma unroll
      for (int g2 = 0; g2 < NG; ++g2) {   // the tile is loaded once for all groups of the CTA
        const double* v = vs + g2 * vsz;
        const double b0 = v[(J * 8 + (lane & 3)) * 8 + (lane >> 2)], b1 = v[(J * 8 + 4 + (lane & 3)) * 8 + (lane >> 2)];
        dmma(c[g2][0], c[g2][1], a0, b0);
        dmma(c[g2][0], c[g2][1], a1, b1);
        dmma(c[g2][2], c[g2][3], m0, b0);
        dmma(c[g2][2], c[g2][3], m1, b1);
      }
    }
#pragma unroll
    for (int g2 = 0; g2 < NG; ++g2) {
      const int grp = grp0 + g2;
      if (!ROWW) cta_reduce<4, KW>(c[g2], red);
      if ((ROWW || warp == 0) && grp < g.G) {
        double* wsb = a.ws + (size_t)(b * g.G + grp) * g.per_bg;
        double* Rr = wsb + off_R(g, D);
        const double* v = vs + g2 * vsz;
        const int i = I * 8 + (lane >> 2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = 2 * (lane & 3) + h, r = grp * kCh + ch;
          double rv = 0.0, t1 = 0.0;
          if (i < n && r < R) {
            const double* xp = a.X + (((size_t)b * R + r) * n + i) * D;
            double x[D], fo[D];
#pragma unroll
            for (int dd = 0; dd < D; ++dd) x[dd] = xp[dd];
            M::f(x, ths + (g2 * kCh + ch) * P, fo);
            double fd = 0.0;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) fd = dd == d ? fo[dd] : fd;
            rv = fd - c[g2][2 + h];
            a.gX[(((size_t)b * R + r) * n + i) * D + d] = 2.0 * c[g2][h];   // scratch: 2 u, finished in pass 3
            t1 = v[i * 8 + ch] * c[g2][h];
          }
          Rr[((size_t)d * g.np + i) * kCh + ch] = rv;
          t1 = rows_sum(t1);
          if (lane < 4) wsb[off_t1(g, D) + ((size_t)d * g.nblk + I) * kCh + ch] = t1;
        }
      }
    }
  }
}

// ---- pass 2: q = S_K r, t2 ---------------------------------------------------------------------------------------
template <bool ROWW, int NG, int KW>
__global__ void __launch_bounds__(32 * KW) wide_pass2(Args a, int D) {
  extern __shared__ double sm[];
  constexpr int kT = 32 * KW;
  const Geo& g = a.g;
  const size_t vsz = (size_t)g.np * kCh;
  double* vs = sm;                          // [NG][np][8]
  double* red = vs + NG * vsz;
  const int grp0 = (blockIdx.x % g.GS) * NG, chunk = blockIdx.x / g.GS, d = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int g2 = 0; g2 < NG; ++g2) {
    const bool live = grp0 + g2 < g.G;
    const double* Rr = a.ws + (size_t)(b * g.G + (live ? grp0 + g2 : grp0)) * g.per_bg + off_R(g, D) + (size_t)d * vsz;
    for (int e = tid; e < g.np * kCh; e += kT) vs[g2 * vsz + e] = live ? Rr[e] : 0.0;
  }
  const double* matK = static_cast<const double*>(a.pb.packed) + ((size_t)(b * D + d) * 3 + 2) * g.np * g.np;
  __syncthreads();
  MAGI_WIDE_ROWS(I) {
    int lo, hi;
    jrange(g, I, lo, hi);
    double c[NG][2];
#pragma unroll
    for (int g2 = 0; g2 < NG; ++g2) c[g2][0] = c[g2][1] = 0.0;
#pragma unroll 8
    for (int J = ROWW ? lo : lo + warp; J <= hi; J += ROWW ? 1 : KW) {
      const size_t t = ((size_t)I * g.nblk + J) * 64;
      const double a0 = matK[t + magi_tile_pos(lane >> 2, lane & 3)], a1 = matK[t + magi_tile_pos(lane >> 2, 4 + (lane & 3))];
#pragma unroll
      for (int g2 = 0; g2 < NG; ++g2) {
        const double* v = vs + g2 * vsz;
        dmma(c[g2][0], c[g2][1], a0, v[(J * 8 + (lane & 3)) * 8 + (lane >> 2)]);
        dmma(c[g2][0], c[g2][1], a1, v[(J * 8 + 4 + (lane & 3)) * 8 + (lane >> 2)]);
      }
    }
#pragma unroll
    for (int g2 = 0; g2 < NG; ++g2) {
      const int grp = grp0 + g2;
      if (!ROWW) cta_reduce<2, KW>(c[g2], red);
      if ((ROWW || warp == 0) && grp < g.G) {
        double* wsb = a.ws + (size_t)(b * g.G + grp) * g.per_bg;
        double* Q = wsb + off_Q(g, D) + (size_t)d * vsz;
        const double* v = vs + g2 * vsz;
        const int i = I * 8 + (lane >> 2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = 2 * (lane & 3) + h;
          Q[(size_t)i * kCh + ch] = c[g2][h];      // rows >= n: zero tiles times zero r -> 0
          const double t2 = rows_sum(v[i * 8 + ch] * c[g2][h]);
          if (lane < 4) wsb[off_t2(g, D) + ((size_t)d * g.nblk + I) * kCh + ch] = t2;
        }
      }
    }
  }
}

// ---- pass 3: m^T q and the point-wise assembly of dX; SSE and d/d theta partial sums -------------------------------
template <class M, bool ROWW, int NG, int KW>
__global__ void __launch_bounds__(32 * KW) wide_pass3(Args a) {
  extern __shared__ double sm[];
  constexpr int D = M::D, P = M::P, kT = 32 * KW;
  const Geo& g = a.g;
  const size_t vsz = (size_t)g.np * kCh;
  double* vs = sm;                          // [NG][np][8]
  double* red = vs + NG * vsz;
  double* ths = red + KW * 4 * 32;          // [NG][8][P]
  // the chain groups of a dataset are adjacent in launch order: they read the same matrix rows at the same time (L2 hits)
  const int grp0 = (blockIdx.x % g.GS) * NG, chunk = blockIdx.x / g.GS, d = blockIdx.y, b = blockIdx.z;
  const int n = a.pb.n, R = a.pb.R, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double bt[NG][2], isig2[NG][2];
#pragma unroll
  for (int g2 = 0; g2 < NG; ++g2) {
    const bool live = grp0 + g2 < g.G;
    const double* Qd = a.ws + (size_t)(b * g.G + (live ? grp0 + g2 : grp0)) * g.per_bg + off_Q(g, D) + (size_t)d * vsz;
    for (int e = tid; e < g.np * kCh; e += kT) vs[g2 * vsz + e] = live ? Qd[e] : 0.0;
    for (int e = tid; e < kCh * P; e += kT) {
      const int ch = e / P, k = e % P, r = (grp0 + g2) * kCh + ch;
      ths[g2 * kCh * P + e] = r < R ? magi_softplus(a.th_pre[((size_t)b * R + r) * P + k]) : 1.0;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = (grp0 + g2) * kCh + 2 * (lane & 3) + h;
      bt[g2][h] = r < R ? a.beta_temp[(size_t)b * R + r] : 0.0;
      isig2[g2][h] = r < R ? 1.0 / (magi_softplus(a.sig_pre[((size_t)b * R + r) * D + d]) + a.pb.LB[b * D + d]) : 0.0;
    }
  }
  const double* matM = static_cast<const double*>(a.pb.packed) + ((size_t)(b * D + d) * 3 + 1) * g.np * g.np;
  const double inv_beta = 1.0 / a.pb.beta[b];
  __syncthreads();
  MAGI_WIDE_ROWS(I) {
    int lo, hi;
    jrange(g, I, lo, hi);
    double c[NG][2];
#pragma unroll
    for (int g2 = 0; g2 < NG; ++g2) c[g2][0] = c[g2][1] = 0.0;
#pragma unroll 8
    for (int J = ROWW ? lo : lo + warp; J <= hi; J += ROWW ? 1 : KW) {
      // (m^T)(I, J) = tile (J, I) transposed: A[row][k] = tile[k][row]
      const size_t t = ((size_t)J * g.nblk + I) * 64;
      const double a0 = matM[t + magi_tile_pos(lane & 3, lane >> 2)], a1 = matM[t + magi_tile_pos(4 + (lane & 3), lane >> 2)];
#pragma unroll
      for (int g2 = 0; g2 < NG; ++g2) {
        const double* v = vs + g2 * vsz;
        dmma(c[g2][0], c[g2][1], a0, v[(J * 8 + (lane & 3)) * 8 + (lane >> 2)]);
        dmma(c[g2][0], c[g2][1], a1, v[(J * 8 + 4 + (lane & 3)) * 8 + (lane >> 2)]);
      }
    }
#pragma unroll
    for (int g2 = 0; g2 < NG; ++g2) {
      const int grp = grp0 + g2;
      if (!ROWW) cta_reduce<2, KW>(c[g2], red);
      if ((ROWW || warp == 0) && grp < g.G) {
        double* wsb = a.ws + (size_t)(b * g.G + grp) * g.per_bg;
        const double* Qall = wsb + off_Q(g, D);
        const int i = I * 8 + (lane >> 2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = 2 * (lane & 3) + h, r = grp * kCh + ch;
          double sse = 0.0, vth[P];
#pragma unroll
          for (int k = 0; k < P; ++k) vth[k] = 0.0;
          if (i < n && r < R) {
            const size_t xo = (((size_t)b * R + r) * n + i) * D;
            double x[D], gq[D], vx[D];
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
              x[dd] = a.X[xo + dd];
              gq[dd] = 2.0 * Qall[((size_t)dd * g.np + i) * kCh + ch];
            }
            M::vjp(x, ths + (g2 * kCh + ch) * P, gq, vx, vth);
            double xd = 0.0, vxd = 0.0;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
              xd = dd == d ? x[dd] : xd;
              vxd = dd == d ? vx[dd] : vxd;
            }
            const double prior = a.gX[xo + d] + vxd - 2.0 * c[g2][h];
            const size_t yo = ((size_t)b * n + i) * D + d;
            const double e = a.pb.mask[yo] ? xd - a.pb.y[yo] : 0.0;
            a.gX[xo + d] = bt[g2][h] * -0.5 * (prior * inv_beta + 2.0 * e * isig2[g2][h]);
            sse = e * e;
          }
          sse = rows_sum(sse);
          if (lane < 4) wsb[off_sse(g, D) + ((size_t)d * g.nblk + I) * kCh + ch] = sse;
          if (d == 0) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
              const double tk = rows_sum(vth[k]);
              if (lane < 4) wsb[off_th(g, D) + ((size_t)I * P + k) * kCh + ch] = tk;
            }
          }
        }
      }
    }
  }
}

// ---- final: per chain sums of the partials -> lp, d/d sigma_pre, d/d theta_pre --------------------------------------
// One CTA per (dataset, chain group), one warp per chain: the lanes stride over the per-block-row partial sums and
// combine them by shuffles (fixed order: deterministic).
__global__ void __launch_bounds__(32 * kCh) wide_final(Args a, int D, int P) {   // one warp per chain of a group
  const Geo& g = a.g;
  const int bg = blockIdx.x, ch = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = bg / g.G, r = (bg % g.G) * kCh + ch;
  if (r >= a.pb.R) return;
  const double* wsb = a.ws + (size_t)bg * g.per_bg;
  const double bt = a.beta_temp[(size_t)b * a.pb.R + r], inv_beta = 1.0 / a.pb.beta[b];
  double t12 = 0.0;
  for (int e = lane; e < D * g.nblk; e += 32)
    t12 += wsb[off_t1(g, D) + (size_t)e * kCh + ch] + wsb[off_t2(g, D) + (size_t)e * kCh + ch];
  t12 = magi_warp_sum(t12);
  double t34 = 0.0, logJ = 0.0;
  for (int d = 0; d < D; ++d) {
    double sse = 0.0;
    for (int c = lane; c < g.nblk; c += 32) sse += wsb[off_sse(g, D) + ((size_t)d * g.nblk + c) * kCh + ch];
    sse = magi_warp_sum(sse);
    const double s = a.sig_pre[((size_t)b * a.pb.R + r) * D + d];
    const double sig2 = magi_softplus(s) + a.pb.LB[b * D + d], Nd = a.pb.N_ds[b * D + d];
    t34 += Nd * log(2.0 * M_PI * sig2) + sse / sig2;
    logJ += s - magi_softplus(s);
    const double sg = magi_sigmoid(s);
    if (lane == 0)
      a.gsig[((size_t)b * a.pb.R + r) * D + d] = bt * (-0.5 * (Nd / sig2 - sse / (sig2 * sig2)) * sg + (1.0 - sg));
  }
  for (int k = 0; k < P; ++k) {
    double v = 0.0;
    for (int c = lane; c < g.nblk; c += 32) v += wsb[off_th(g, D) + ((size_t)c * P + k) * kCh + ch];
    v = magi_warp_sum(v);
    const double tau = a.th_pre[((size_t)b * a.pb.R + r) * P + k];
    const double sg = magi_sigmoid(tau);
    logJ += tau - magi_softplus(tau);
    if (lane == 0) a.gth[((size_t)b * a.pb.R + r) * P + k] = bt * (-0.5 * inv_beta * v * sg + (1.0 - sg));
  }
  if (lane == 0) a.lp[(size_t)b * a.pb.R + r] = bt * (-0.5 * (t12 * inv_beta + t34) + logJ);
}

int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <class M, bool ROWW, int NG, int KW>
int launch_wide_t(const Args& a, cudaStream_t st) {
  const Geo& g = a.g;
  const dim3 grid(g.nchunk * g.GS, M::D, a.pb.B);
  if (grid.z > 65535) return MAGI_ERR_UNSUPPORTED;
  const size_t smem = (NG * (size_t)g.np * kCh + KW * 4 * 32 + NG * kCh * M::P) * sizeof(double);
  if (smem > 200 * 1024) return MAGI_ERR_UNSUPPORTED;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(wide_pass1<M, ROWW, NG, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return magi_cuda_status(e);
  if ((e = cudaFuncSetAttribute(wide_pass2<ROWW, NG, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return magi_cuda_status(e);
  if ((e = cudaFuncSetAttribute(wide_pass3<M, ROWW, NG, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return magi_cuda_status(e);
  wide_pass1<M, ROWW, NG, KW><<<grid, 32 * KW, smem, st>>>(a);
  wide_pass2<ROWW, NG, KW><<<grid, 32 * KW, smem, st>>>(a, M::D);
  wide_pass3<M, ROWW, NG, KW><<<grid, 32 * KW, smem, st>>>(a);
  wide_final<<<a.pb.B * g.G, 32 * kCh, 0, st>>>(a, M::D, M::P);
  return magi_cuda_status(cudaGetLastError());
}

template <class M, int NG, int KW>
int launch_wide_k(const Args& a, cudaStream_t st) {
  const char* force = getenv("MAGI_WIDE_ROWW");   // experiment knob: 0 / 1 overrides the choice of work split
  const bool roww = force ? (force[0] == '1' && a.g.rbpc <= KW) : a.g.rbpc == KW;
  return roww ? launch_wide_t<M, true, NG, KW>(a, st) : launch_wide_t<M, false, NG, KW>(a, st);
}

template <class M>
int launch_wide(const Args& a, cudaStream_t st) {
  if (a.g.KW == 24) return launch_wide_k<M, 1, 24>(a, st);
  if (a.g.KW == 16) return launch_wide_k<M, 1, 16>(a, st);
  return a.g.NG == 2 ? launch_wide_k<M, 2, 8>(a, st) : launch_wide_k<M, 1, 8>(a, st);
}

int model_dims(int id, int& D, int& P) {
  switch (id) {
#ifdef MAGI_USER_MODEL_HEADER
    case MAGI_MODEL_USER: D = UserModel::D; P = UserModel::P; return 0;
#else
    case MAGI_MODEL_SEIR3: D = Seir3::D; P = Seir3::P; return 0;
    case MAGI_MODEL_SEIR4: D = Seir4::D; P = Seir4::P; return 0;
    case MAGI_MODEL_SIRW: D = Sirw::D; P = Sirw::P; return 0;
    case MAGI_MODEL_LORENZ96: D = Lorenz96::D; P = Lorenz96::P; return 0;
#endif
    default: return -1;
  }
}

}  // namespace

extern "C" size_t magi_b200_logpost_grad_wide_workspace_bytes(const magi_problem_t* pb) {
  if (!pb || pb->B <= 0 || pb->R <= 0 || pb->n <= 1 || pb->D <= 0) return 0;
  const Geo g = make_geo(pb, sm_count());
  return g.per_bg * (size_t)pb->B * g.G * sizeof(double);
}

extern "C" int magi_b200_logpost_grad_wide(const magi_problem_t* pb, const double* X, const double* sig_pre,
                                           const double* th_pre, const double* beta_temp, double* lp, double* gX,
                                           double* gsig, double* gth, void* ws, size_t ws_bytes,
                                           magi_stream_t stream) {
  if (!pb) return -1;
  int D, P;
  if (model_dims(pb->model_id, D, P) != 0) return MAGI_ERR_UNSUPPORTED;
  if (pb->D != D || pb->P != P || pb->B <= 0 || pb->R <= 0 || pb->n <= 1) return -1;
  if (!pb->packed || !pb->mu || !pb->y || !pb->mask || !pb->N_ds || !pb->beta || !pb->LB) return -1;
  if (!X) return -2;
  if (!sig_pre) return -3;
  if (!th_pre) return -4;
  if (!beta_temp) return -5;
  if (!lp) return -6;
  if (!gX) return -7;
  if (!gsig) return -8;
  if (!gth) return -9;
  Args a;
  a.pb = *pb;
  a.g = make_geo(pb, sm_count());
  if (!ws || ws_bytes < a.g.per_bg * (size_t)pb->B * a.g.G * sizeof(double)) return -10;
  a.X = X; a.sig_pre = sig_pre; a.th_pre = th_pre; a.beta_temp = beta_temp;
  a.lp = lp; a.gX = gX; a.gsig = gsig; a.gth = gth; a.ws = static_cast<double*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (pb->model_id) {
#ifdef MAGI_USER_MODEL_HEADER
    case MAGI_MODEL_USER: return launch_wide<UserModel>(a, st);
#else
    case MAGI_MODEL_SEIR3: return launch_wide<Seir3>(a, st);
    case MAGI_MODEL_SEIR4: return launch_wide<Seir4>(a, st);
    case MAGI_MODEL_SIRW: return launch_wide<Sirw>(a, st);
    case MAGI_MODEL_LORENZ96: return launch_wide<Lorenz96>(a, st);
#endif
    default: return MAGI_ERR_UNSUPPORTED;
  }
}
