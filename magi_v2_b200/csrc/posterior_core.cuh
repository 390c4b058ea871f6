// Device core of the MAGI log-posterior + analytic gradient (replaces magi_v2.py:308-348 and the
// TF reverse-mode gradient TFP's leapfrog takes of it).  One CTA owns one dataset and a group of
// up to 8 of its chains; the dataset's packed matrices sym(C^-1) | m | sym(K^-1) are streamed from
// HBM/L2 once per evaluation and shared by the 8 chains.  Shared by the logpost_grad, leapfrog and
// HMC kernels.
//
// Math, per chain, "base" quantities WITHOUT the temperature factor beta_temp
// (lp = beta_temp * L, grad lp = beta_temp * grad L):
//   u_d = S_C,d xc_d          t1 = sum_d xc_d . u_d            S_C = (C^-1 + C^-T)/2
//   r_d = f_d(X,theta) - m_d xc_d
//   q_d = S_K,d r_d           t2 = sum_d r_d . q_d ;  g_d = 2 q_d
//   dL/dX[:,d] = -1/2 { [2 u_d + sum_d' J[d',d] g_d' - m_d^T g_d] / beta + 2 M (X - y) / sigma_d^2 }
#pragma once
#include "common.cuh"
#include "ode_models.cuh"

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kCh = MAGI_CHAINS_PER_CTA;  // 8

// Scratch arrays of one CTA.  Vector arrays are [D][8][np] (chain-major, grid index fastest) so
// that lanes walking the grid index hit consecutive shared-memory banks.
template <class M>
struct Scratch {
  int n, np;
  double *Xc, *FR, *G, *GX;  // [D][8][np]   centred state, f -> residual r, g = 2 q, gradient
  double* PX;                // [D][8][np]   momentum (HMC kernels only; else nullptr)
  double* W;                 // [8][np]      m_d xc_d
  double *Y, *MK;            // [D][np]      observations / mask as 0.0 or 1.0
  double *tau, *th, *sgt;    // [P][8]       theta_pre, softplus, sigmoid
  double *s, *sig2, *sgs;    // [D][8]       sigma_pre, softplus + LB, sigmoid
  double *ptau, *ps;         // [P][8], [D][8] momenta of the small state parts
  double *gtau, *gs;         // [P][8], [D][8] base gradient w.r.t. tau, s
  double* L;                 // [8]          base log-posterior
  double *mu, *Nd, *LB;      // [D]
  double* wpart;             // [kWarps][NRED]
  double* ctl;               // [kCtl][8]    sampler control values (step size, energies, saved small state)
  static constexpr int kCtl = 16 + 2 * M::P + 2 * M::D;
  static constexpr int NRED = 2 + M::D + M::P;

  __host__ __device__ static size_t big_elems(int np, bool with_momentum) {
    return (size_t)(with_momentum ? 5 : 4) * M::D * kCh * np + (size_t)kCh * np;
  }
  __host__ __device__ static size_t small_elems(int np) {
    const size_t e = (size_t)2 * M::D * np + (size_t)kCh * (5 * M::P + 5 * M::D + 1 + kCtl) + 3 * M::D +
                     (size_t)kWarps * NRED;
    return (e + 1) & ~(size_t)1;  // keep the big arrays behind it 16-byte aligned (double2 loads)
  }
  __device__ void carve(double* big, double* small, int n_, int np_, bool with_momentum) {
    n = n_;
    np = np_;
    const size_t v = (size_t)M::D * kCh * np;
    Xc = big; FR = Xc + v; G = FR + v; GX = G + v;
    double* p = GX + v;
    PX = nullptr;
    if (with_momentum) { PX = p; p += v; }
    W = p;
    p = small;
    Y = p; p += M::D * np;
    MK = p; p += M::D * np;
    tau = p; p += M::P * kCh; th = p; p += M::P * kCh; sgt = p; p += M::P * kCh;
    ptau = p; p += M::P * kCh; gtau = p; p += M::P * kCh;
    s = p; p += M::D * kCh; sig2 = p; p += M::D * kCh; sgs = p; p += M::D * kCh;
    ps = p; p += M::D * kCh; gs = p; p += M::D * kCh;
    L = p; p += kCh;
    mu = p; p += M::D; Nd = p; p += M::D; LB = p; p += M::D;
    wpart = p; p += kWarps * NRED;
    ctl = p;
  }
  __device__ __forceinline__ size_t vix(int d, int r, int j) const { return ((size_t)d * kCh + r) * np + j; }
};

// ---- reduce 16 per-lane values across the warp: butterfly that halves the value count per round.
// On return lane L holds in v[0] the full sum of value index
//   ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
__device__ __forceinline__ void warp_reduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool up = lane & 16;
    const double send = up ? v[k] : v[k + 8];
    const double keep = up ? v[k + 8] : v[k];
    v[k] = keep + magi_shfl_xor(send, 16);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up = lane & 8;
    const double send = up ? v[k] : v[k + 4];
    const double keep = up ? v[k + 4] : v[k];
    v[k] = keep + magi_shfl_xor(send, 8);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 4;
    const double send = up ? v[k] : v[k + 2];
    const double keep = up ? v[k + 2] : v[k];
    v[k] = keep + magi_shfl_xor(send, 4);
  }
  {
    const bool up = lane & 2;
    const double send = up ? v[0] : v[1];
    const double keep = up ? v[1] : v[0];
    v[0] = keep + magi_shfl_xor(send, 2);
  }
  v[0] += magi_shfl_xor(v[0], 1);
}

// y1[r][i] = scale1 * sum_j A1[i][j] x[r][j],  y2[r][i] = scale2 * sum_j A2[i][j] x[r][j]   (i < n)
// Warp per row; lanes stride the row with 16-byte loads (coalesced 512 B per warp instruction).
// A2 == nullptr: single matrix.
__device__ __forceinline__ void matvec_rows(const double* __restrict__ A1, const double* __restrict__ A2,
                                            const double* x, double* y1, double* y2, double scale1,
                                            double scale2, int n, int np) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npair = np >> 1;
  for (int i = warp; i < n; i += kWarps) {
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    const double2* r1 = reinterpret_cast<const double2*>(A1 + (size_t)i * np);
    const double2* r2 = A2 ? reinterpret_cast<const double2*>(A2 + (size_t)i * np) : nullptr;
    for (int jp = lane; jp < npair; jp += 32) {
      const double2 a1 = __ldg(r1 + jp);
      double2 a2 = make_double2(0.0, 0.0);
      if (r2) a2 = __ldg(r2 + jp);
#pragma unroll
      for (int r = 0; r < kCh; ++r) {
        const double2 xv = *reinterpret_cast<const double2*>(x + (size_t)r * np + 2 * jp);
        acc[r] = fma(a1.x, xv.x, acc[r]);
        acc[r] = fma(a1.y, xv.y, acc[r]);
        acc[8 + r] = fma(a2.x, xv.x, acc[8 + r]);
        acc[8 + r] = fma(a2.y, xv.y, acc[8 + r]);
      }
    }
    warp_reduce16(acc, lane);
    if ((lane & 1) == 0) {
      const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      if (idx < 8) y1[(size_t)idx * np + i] = scale1 * acc[0];
      else if (y2) y2[(size_t)(idx - 8) * np + i] = scale2 * acc[0];
    }
  }
}

// y[r][j] -= sum_i A[i][j] g[r][i]   (transposed product, A row-major [n][np]).
// Thread owns a column pair and a slice of the rows; slices are combined in a fixed order
// (deterministic, no atomics).  Contains __syncthreads(): call from all threads.
__device__ __forceinline__ void matvec_cols_sub(const double* __restrict__ A, const double* g, double* y, int n,
                                                int np) {
  const int tid = threadIdx.x;
  const int npair = np >> 1;
  const int nsplit = npair < kThreads ? kThreads / npair : 1;
  const int s = npair < kThreads ? tid / npair : 0;
  const int jp0 = tid - s * npair;
  if (nsplit == 1) {
    for (int jp = tid; jp < npair; jp += kThreads) {
      double acc[2][kCh];
#pragma unroll
      for (int r = 0; r < kCh; ++r) acc[0][r] = acc[1][r] = 0.0;
      for (int i = 0; i < n; ++i) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(A + (size_t)i * np) + jp);
#pragma unroll
        for (int r = 0; r < kCh; ++r) {
          const double gv = g[(size_t)r * np + i];
          acc[0][r] = fma(a.x, gv, acc[0][r]);
          acc[1][r] = fma(a.y, gv, acc[1][r]);
        }
      }
#pragma unroll
      for (int r = 0; r < kCh; ++r) {
        y[(size_t)r * np + 2 * jp] -= acc[0][r];
        y[(size_t)r * np + 2 * jp + 1] -= acc[1][r];
      }
    }
    __syncthreads();
    return;
  }
  double acc[2][kCh];
#pragma unroll
  for (int r = 0; r < kCh; ++r) acc[0][r] = acc[1][r] = 0.0;
  if (s < nsplit) {
    for (int i = s; i < n; i += nsplit) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(A + (size_t)i * np) + jp0);
#pragma unroll
      for (int r = 0; r < kCh; ++r) {
        const double gv = g[(size_t)r * np + i];
        acc[0][r] = fma(a.x, gv, acc[0][r]);
        acc[1][r] = fma(a.y, gv, acc[1][r]);
      }
    }
  }
  for (int ss = 0; ss < nsplit; ++ss) {
    if (s == ss) {
#pragma unroll
      for (int r = 0; r < kCh; ++r) {
        y[(size_t)r * np + 2 * jp0] -= acc[0][r];
        y[(size_t)r * np + 2 * jp0 + 1] -= acc[1][r];
      }
    }
    __syncthreads();
  }
}

// Per-chain transforms of the small state parts (magi_v2.py:318-319).  Threads 0..7.
template <class M>
__device__ __forceinline__ void chain_scalars(const Scratch<M>& S) {
  const int r = threadIdx.x;
  if (r < kCh) {
#pragma unroll
    for (int k = 0; k < M::P; ++k) {
      const double t = S.tau[k * kCh + r];
      S.th[k * kCh + r] = magi_softplus(t);
      S.sgt[k * kCh + r] = magi_sigmoid(t);
    }
#pragma unroll
    for (int d = 0; d < M::D; ++d) {
      const double z = S.s[d * kCh + r];
      S.sig2[d * kCh + r] = magi_softplus(z) + S.LB[d];
      S.sgs[d * kCh + r] = magi_sigmoid(z);
    }
  }
}

// Evaluate base log-posterior L and its gradient at the state held in scratch
// (S.Xc centred trajectories, S.tau, S.s).  Results: S.L[8], S.GX (dL/dX), S.gtau, S.gs.
// `mats` = packed matrices of this dataset: [D][3][np][np].  All threads must call.
template <class M>
__device__ void eval_logpost_grad(const Scratch<M>& S, const double* __restrict__ mats, double inv_beta) {
  constexpr int D = M::D, P = M::P, NRED = Scratch<M>::NRED;
  const int n = S.n, np = S.np;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = tid >> 6, l = tid & 63;  // pointwise phases: 64 threads per chain
  const size_t msz = (size_t)np * np;

  chain_scalars(S);
  __syncthreads();

  double th[P];
#pragma unroll
  for (int k = 0; k < P; ++k) th[k] = S.th[k * kCh + r];

  // f(X, theta) at every grid point
  for (int j = l; j < np; j += 64) {
    double x[D], f[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = S.Xc[S.vix(d, r, j)] + S.mu[d];
    M::f(x, th, f);
#pragma unroll
    for (int d = 0; d < D; ++d) S.FR[S.vix(d, r, j)] = j < n ? f[d] : 0.0;
  }
  __syncthreads();

  double t1 = 0.0;
  for (int d = 0; d < D; ++d) {
    const double* SC = mats + (size_t)(3 * d + 0) * msz;
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    double* xc = S.Xc + S.vix(d, 0, 0);
    double* fr = S.FR + S.vix(d, 0, 0);
    double* g = S.G + S.vix(d, 0, 0);
    double* gx = S.GX + S.vix(d, 0, 0);
    // gx = 2 S_C xc ; W = m xc
    matvec_rows(SC, Mm, xc, gx, S.W, 2.0, 1.0, n, np);
    __syncthreads();
    for (int j = l; j < n; j += 64) {
      const size_t a = (size_t)r * np + j;
      t1 = fma(xc[a], 0.5 * gx[a], t1);
      fr[a] -= S.W[a];
    }
    __syncthreads();
    // g = 2 S_K r
    matvec_rows(SK, nullptr, fr, g, nullptr, 2.0, 0.0, n, np);
    __syncthreads();
    // gx -= m^T g
    matvec_cols_sub(Mm, g, gx, n, np);  // ends with __syncthreads()
  }

  // pointwise: ODE Jacobian terms, likelihood, assemble gradient
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) red[k] = 0.0;
  red[0] = t1;
  double isig2[D];
#pragma unroll
  for (int d = 0; d < D; ++d) isig2[d] = 1.0 / S.sig2[d * kCh + r];
  for (int j = l; j < n; j += 64) {
    double x[D], g[D], vx[D], vth[P];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      x[d] = S.Xc[S.vix(d, r, j)] + S.mu[d];
      g[d] = S.G[S.vix(d, r, j)];
      red[1] = fma(S.FR[S.vix(d, r, j)], 0.5 * g[d], red[1]);
    }
    M::vjp(x, th, g, vx, vth);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double e = S.MK[d * np + j] != 0.0 ? x[d] - S.Y[d * np + j] : 0.0;
      red[2 + d] = fma(e, e, red[2 + d]);
      const size_t a = S.vix(d, r, j);
      S.GX[a] = -0.5 * ((S.GX[a] + vx[d]) * inv_beta + 2.0 * e * isig2[d]);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) red[2 + D + k] += vth[k];
  }
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    const double v = magi_warp_sum(red[k]);
    if (lane == 0) S.wpart[warp * NRED + k] = v;
  }
  __syncthreads();
  if (tid < kCh) {
    const int c = tid;  // chain; its two warps are 2c and 2c+1
    double tot[NRED];
#pragma unroll
    for (int k = 0; k < NRED; ++k) tot[k] = S.wpart[(2 * c) * NRED + k] + S.wpart[(2 * c + 1) * NRED + k];
    double t3 = 0.0, t4 = 0.0, lj = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double s2 = S.sig2[d * kCh + c], sg = S.sgs[d * kCh + c], z = S.s[d * kCh + c];
      t3 += S.Nd[d] * log(2.0 * M_PI * s2);
      t4 += tot[2 + d] / s2;
      lj += z - magi_softplus(z);
      S.gs[d * kCh + c] = -0.5 * (S.Nd[d] / s2 - tot[2 + d] / (s2 * s2)) * sg + (1.0 - sg);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double t = S.tau[k * kCh + c], sg = S.sgt[k * kCh + c];
      lj += t - magi_softplus(t);
      S.gtau[k * kCh + c] = -0.5 * inv_beta * tot[2 + D + k] * sg + (1.0 - sg);
    }
    S.L[c] = -0.5 * ((inv_beta * (tot[0] + tot[1])) + (t3 + t4)) + lj;
  }
  __syncthreads();
}

// Load the per-dataset constants and zero the padding of the vector arrays.
template <class M>
__device__ void load_dataset(const Scratch<M>& S, const magi_problem_t& pb, int b) {
  constexpr int D = M::D;
  const int n = S.n, np = S.np, tid = threadIdx.x;
  if (tid < D) {
    S.mu[tid] = pb.mu[(size_t)b * D + tid];
    S.Nd[tid] = pb.N_ds[(size_t)b * D + tid];
    S.LB[tid] = pb.LB[(size_t)b * D + tid];
  }
  for (int e = tid; e < D * np; e += kThreads) {
    const int d = e / np, j = e - d * np;
    double yv = 0.0, mk = 0.0;
    if (j < n) {
      const size_t a = ((size_t)b * n + j) * D + d;
      mk = pb.mask[a] ? 1.0 : 0.0;
      yv = mk != 0.0 ? pb.y[a] : 0.0;
    }
    S.Y[e] = yv;
    S.MK[e] = mk;
  }
}

// Load chain state (reference layout X[n][D] per chain) into scratch, centring X.
// Chains >= nr and grid indices >= n are zero-filled.  Needs S.mu loaded (sync before).
template <class M>
__device__ void load_state(const Scratch<M>& S, const double* X, const double* sig_pre, const double* th_pre,
                           size_t chain0, int nr) {
  constexpr int D = M::D, P = M::P;
  const int n = S.n, np = S.np, tid = threadIdx.x;
  const size_t tot = (size_t)D * kCh * np;
  for (size_t e = tid; e < tot; e += kThreads) S.Xc[e] = 0.0;
  __syncthreads();
  const int per = n * D;
  for (int e = tid; e < nr * per; e += kThreads) {
    const int r = e / per, rem = e - r * per;
    const int j = rem / D, d = rem - j * D;
    S.Xc[S.vix(d, r, j)] = X[(chain0 + r) * per + rem] - S.mu[d];
  }
  if (tid < kCh * D) {
    const int r = tid / D, d = tid - r * D;
    S.s[d * kCh + r] = r < nr ? sig_pre[(chain0 + r) * D + d] : 0.0;
  }
  if (tid >= 64 && tid < 64 + kCh * P) {
    const int t = tid - 64, r = t / P, k = t - r * P;
    S.tau[k * kCh + r] = r < nr ? th_pre[(chain0 + r) * P + k] : 0.0;
  }
}
