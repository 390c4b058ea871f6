// Device core of the MAGI log-posterior + analytic gradient (replaces magi_v2.py:308-348 and the
// TF reverse-mode gradient TFP's leapfrog takes of it).  One CTA owns one dataset and a group of
// up to 8 of its chains.  The dataset's packed matrices sym(C^-1) | m | sym(K^-1) are streamed from
// HBM/L2 once per evaluation and shared by the 8 chains: with 8 chains per matrix the four batched
// mat-vecs per component become the dense contractions  [n x n] . [n x 8], which run on the FP64
// tensor cores (mma.sync.m8n8k4.f64 -> DMMA) with N = 8 = chains.  Matrix fragments go straight
// from global memory to registers (software-pipelined 128-bit loads, prefetched across phase
// boundaries); the chain vectors are the B operand and live in shared memory.
// Shared by the logpost_grad, leapfrog and HMC kernels.
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

constexpr int kMaxWarps = 21;               // 672 threads: 96 registers per thread, one CTA per SM
constexpr int kMaxThreads = kMaxWarps * 32;
constexpr int kMinWarps = 8;                // pointwise phases map warp w to chain w % 8
constexpr int kCh = MAGI_CHAINS_PER_CTA;    // 8
constexpr int kU = 7;                       // 8-column steps per register batch (7 x 16 B per lane in flight)

// chain stride of the shared-memory vector arrays: np + 2 doubles, so that the 128-bit B-fragment
// loads of a quarter warp (2 chains x 4 column pairs) fall into 32 distinct banks
__host__ __device__ static inline int magi_chain_stride(int np) { return np + 2; }

// Scratch arrays of one CTA.  Vector arrays are [D][8][ns] (chain-major, grid index fastest).
template <class M>
struct Scratch {
  int n, np, ns;
  double *Xc, *FG, *GX;      // [D][8][ns]   centred state | f -> residual r -> g = 2 S_K r | gradient
  double* PX;                // [D][8][ns]   momentum (HMC kernels only; else nullptr)
  double* W;                 // [8][ns]      m_d xc_d, then g_d
  double *Y, *MK;            // [D][np]      observations / mask as 0.0 or 1.0
  double *tau, *th, *sgt;    // [P][8]       theta_pre, softplus, sigmoid
  double *s, *sig2, *sgs;    // [D][8]       sigma_pre, softplus + LB, sigmoid
  double *ptau, *ps;         // [P][8], [D][8] momenta of the small state parts
  double *gtau, *gs;         // [P][8], [D][8] base gradient w.r.t. tau, s
  double* L;                 // [8]          base log-posterior
  double *mu, *Nd, *LB;      // [D]
  double* wpart;             // [kMaxWarps][NRED]
  double* ctl;               // [kCtl][8]    sampler control values (step size, energies, saved small state)
  static constexpr int kCtl = 16 + 2 * M::P + 2 * M::D;
  static constexpr int NRED = 2 + M::D + M::P;

  __host__ __device__ static size_t big_elems(int np, bool with_momentum) {
    return (size_t)(with_momentum ? 4 : 3) * M::D * kCh * magi_chain_stride(np) + (size_t)kCh * magi_chain_stride(np);
  }
  __host__ __device__ static size_t small_elems(int np) {
    const size_t e = (size_t)2 * M::D * np + (size_t)kCh * (5 * M::P + 5 * M::D + 1 + kCtl) + 3 * M::D +
                     (size_t)kMaxWarps * NRED;
    return (e + 1) & ~(size_t)1;  // keep the big arrays behind it 16-byte aligned (128-bit loads)
  }
  __device__ void carve(double* big, double* small, int n_, int np_, bool with_momentum) {
    n = n_;
    np = np_;
    ns = magi_chain_stride(np_);
    const size_t v = (size_t)M::D * kCh * ns;
    Xc = big; FG = Xc + v; GX = FG + v;
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
    wpart = p; p += kMaxWarps * NRED;
    ctl = p;
  }
  __device__ __forceinline__ size_t vix(int d, int r, int j) const { return ((size_t)d * kCh + r) * ns + j; }
  __device__ __forceinline__ size_t vsize() const { return (size_t)M::D * kCh * ns; }
};

// Pointwise phases: warp w serves chain w % 8; the warps of a chain interleave over the grid index.
struct ChainMap {
  int r, j0, jstride, nwr;
};
__device__ __forceinline__ ChainMap chain_map() {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  ChainMap cm;
  cm.r = w & 7;
  cm.nwr = (nw - cm.r + 7) >> 3;  // warps serving chain r (blockDim >= 8 warps)
  cm.j0 = (w >> 3) * 32 + lane;
  cm.jstride = 32 * cm.nwr;
  return cm;
}

// ---- FP64 tensor-core contraction ------------------------------------------------------------
// D[8x8] += A[8x4] B[4x8]; lane = 4*g + c holds a = A[g][c], b = B[c][g], c0/c1 = D[g][2c], D[g][2c+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// A matrix stream as seen by one lane.  Forward (y = A x, 8 output rows i0..i0+7): the lane walks row
// i0 + g in 8-column steps, 16 B at columns 8s + 2c, 2c+1.  Transposed (y = A^T x, 8 output columns
// j0..j0+7): column j0 + g, rows 8s + 2c and 8s + 2c + 1.  Either way step s contributes two k=4
// contractions whose k-index c maps to vector element 8s + 2c (+1), i.e. the B fragments are the
// 16 B at x[chain g][8s + 2c].
struct MatStream {
  const double* p;  // lane base pointer
  int tr;           // 0 forward, 1 transposed
};
__device__ __forceinline__ MatStream stream_fwd(const double* A, int np, int blk, int lane) {
  return MatStream{A + (size_t)(blk * 8 + (lane >> 2)) * np + 2 * (lane & 3), 0};
}
__device__ __forceinline__ MatStream stream_tr(const double* A, int np, int blk, int lane) {
  return MatStream{A + (size_t)(2 * (lane & 3)) * np + blk * 8 + (lane >> 2), 1};
}

__device__ __forceinline__ double2 load_step(const MatStream st, int s, int nk8, int np) {
  double2 v = make_double2(0.0, 0.0);
  if (s < nk8) {
    if (st.tr == 0) {
      v = __ldg(reinterpret_cast<const double2*>(st.p + 8 * s));
    } else {
      v.x = __ldg(st.p + (size_t)(8 * s) * np);
      v.y = __ldg(st.p + (size_t)(8 * s + 1) * np);
    }
  }
  return v;
}

__device__ __forceinline__ void load_batch(double2 (&a)[kU], const MatStream st, int nk8, int np) {
#pragma unroll
  for (int u = 0; u < kU; ++u) a[u] = load_step(st, u, nk8, np);
}

// One 8-row (or 8-column) block of a contraction with the 8 chain vectors x[8][ns]: a rolling
// register pipeline.  On entry a[u] holds step u of `cur` (u < kU); every register is refilled with
// step s + kU as soon as step s has been consumed, so kU 16-byte loads per lane stay in flight; at
// the end of the block the refills switch to the warp's NEXT task `nxt` (possibly in a later phase,
// behind a __syncthreads), whose first kU steps are therefore already in flight when it starts.
// Result: c0, c1 = y[chain 2c], y[chain 2c+1] at block element g (lane = 4g + c).
__device__ __forceinline__ void mma_task(double2 (&a)[kU], const MatStream cur, const MatStream nxt, bool has_next,
                                         const double* x, int ns, int nk8, int np, double& c0, double& c1) {
  const int lane = threadIdx.x & 31;
  const double* bp = x + (size_t)(lane >> 2) * ns + 2 * (lane & 3);
  const int nk8p = ((nk8 + kU - 1) / kU) * kU;  // steps rounded up to whole batches (extra ones are empty)
  const int nk8n = has_next ? nk8 : 0;
  double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
  for (int s0 = 0; s0 < nk8p; s0 += kU) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int s = s0 + u;
      if (s < nk8) {
        const double2 b = *reinterpret_cast<const double2*>(bp + 8 * s);
        dmma(e0, e1, a[u].x, b.x);
        dmma(o0, o1, a[u].y, b.y);
      }
      const int sn = s + kU;
      a[u] = sn < nk8p ? load_step(cur, sn, nk8, np) : load_step(nxt, sn - nk8p, nk8n, np);
    }
  }
  c0 = e0 + o0;
  c1 = e1 + o1;
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
  const int n = S.n, np = S.np, ns = S.ns;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const ChainMap cm = chain_map();
  const int r = cm.r;
  const size_t msz = (size_t)np * np;
  const int nblk = np >> 3;  // 8-row blocks = 8-column steps
  const int g = lane >> 2, c2 = 2 * (lane & 3);

  chain_scalars(S);
  // first matrix batch of this warp goes in flight before anything else
  double2 a[kU];
  const bool active = warp < nblk;
  if (active) load_batch(a, stream_fwd(mats, np, warp, lane), nblk, np);
  __syncthreads();

  double th[P];
#pragma unroll
  for (int k = 0; k < P; ++k) th[k] = S.th[k * kCh + r];

  // f(X, theta) at every grid point (padding stays zero: it is a B operand of the contractions)
  for (int j = cm.j0; j < ns; j += cm.jstride) {
    double x[D], f[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = S.Xc[S.vix(d, r, j)] + S.mu[d];
    M::f(x, th, f);
#pragma unroll
    for (int d = 0; d < D; ++d) S.FG[S.vix(d, r, j)] = j < n ? f[d] : 0.0;
  }
  __syncthreads();

  double t1 = 0.0, t2 = 0.0;
  for (int d = 0; d < D; ++d) {
    const double* SC = mats + (size_t)(3 * d + 0) * msz;
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    double* xc = S.Xc + S.vix(d, 0, 0);
    double* fg = S.FG + S.vix(d, 0, 0);
    double* gx = S.GX + S.vix(d, 0, 0);
    // pass 1: gx = 2 S_C xc ; W = m xc
    for (int blk = warp; blk < nblk; blk += nw) {
      double c0, c1;
      const MatStream sm = stream_fwd(Mm, np, blk, lane);
      mma_task(a, stream_fwd(SC, np, blk, lane), sm, true, xc, ns, nblk, np, c0, c1);
      gx[(size_t)c2 * ns + blk * 8 + g] = 2.0 * c0;
      gx[(size_t)(c2 + 1) * ns + blk * 8 + g] = 2.0 * c1;
      const bool more = blk + nw < nblk;
      const MatStream nx = more ? stream_fwd(SC, np, blk + nw, lane) : stream_fwd(SK, np, warp, lane);
      mma_task(a, sm, nx, true, xc, ns, nblk, np, c0, c1);
      S.W[(size_t)c2 * ns + blk * 8 + g] = c0;
      S.W[(size_t)(c2 + 1) * ns + blk * 8 + g] = c1;
    }
    __syncthreads();
    for (int j = cm.j0; j < n; j += cm.jstride) {
      const size_t e = (size_t)r * ns + j;
      t1 = fma(xc[e], 0.5 * gx[e], t1);
      fg[e] -= S.W[e];
    }
    __syncthreads();
    // pass 2: W = g = 2 S_K r
    for (int blk = warp; blk < nblk; blk += nw) {
      double c0, c1;
      const bool more = blk + nw < nblk;
      const MatStream nx = more ? stream_fwd(SK, np, blk + nw, lane) : stream_tr(Mm, np, warp, lane);
      mma_task(a, stream_fwd(SK, np, blk, lane), nx, true, fg, ns, nblk, np, c0, c1);
      S.W[(size_t)c2 * ns + blk * 8 + g] = 2.0 * c0;
      S.W[(size_t)(c2 + 1) * ns + blk * 8 + g] = 2.0 * c1;
    }
    __syncthreads();
    // pass 3: gx -= m^T g   (second read of m: an L2 hit)
    for (int blk = warp; blk < nblk; blk += nw) {
      double c0, c1;
      const bool more = blk + nw < nblk;
      const bool has = more || d + 1 < D;
      const MatStream nx = more ? stream_tr(Mm, np, blk + nw, lane)
                                : stream_fwd(mats + (size_t)(3 * (d + 1)) * msz, np, warp, lane);
      mma_task(a, stream_tr(Mm, np, blk, lane), nx, has, S.W, ns, nblk, np, c0, c1);
      gx[(size_t)c2 * ns + blk * 8 + g] -= c0;
      gx[(size_t)(c2 + 1) * ns + blk * 8 + g] -= c1;
    }
    // t2 partial, and g_d takes the place of r_d (neither touches what pass 3 reads or writes)
    for (int j = cm.j0; j < n; j += cm.jstride) {
      const size_t e = (size_t)r * ns + j;
      const double gv = S.W[e];
      t2 = fma(fg[e], 0.5 * gv, t2);
      fg[e] = gv;
    }
    __syncthreads();
  }

  // pointwise: ODE Jacobian terms, likelihood, assemble gradient
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) red[k] = 0.0;
  red[0] = t1;
  red[1] = t2;
  double isig2[D];
#pragma unroll
  for (int d = 0; d < D; ++d) isig2[d] = 1.0 / S.sig2[d * kCh + r];
  for (int j = cm.j0; j < n; j += cm.jstride) {
    double x[D], gg[D], vx[D], vth[P];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      x[d] = S.Xc[S.vix(d, r, j)] + S.mu[d];
      gg[d] = S.FG[S.vix(d, r, j)];
    }
    M::vjp(x, th, gg, vx, vth);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double e = S.MK[d * np + j] != 0.0 ? x[d] - S.Y[d * np + j] : 0.0;
      red[2 + d] = fma(e, e, red[2 + d]);
      const size_t ai = S.vix(d, r, j);
      S.GX[ai] = -0.5 * ((S.GX[ai] + vx[d]) * inv_beta + 2.0 * e * isig2[d]);
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
    const int c = tid;  // chain; its warps are c, c + 8, c + 16 (fixed summation order)
    double tot[NRED];
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      double v = 0.0;
      for (int w = c; w < nw; w += 8) v += S.wpart[w * NRED + k];
      tot[k] = v;
    }
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

// Load the per-dataset constants.
template <class M>
__device__ void load_dataset(const Scratch<M>& S, const magi_problem_t& pb, int b) {
  constexpr int D = M::D;
  const int n = S.n, np = S.np, tid = threadIdx.x;
  if (tid < D) {
    S.mu[tid] = pb.mu[(size_t)b * D + tid];
    S.Nd[tid] = pb.N_ds[(size_t)b * D + tid];
    S.LB[tid] = pb.LB[(size_t)b * D + tid];
  }
  for (int e = tid; e < D * np; e += blockDim.x) {
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
  const int n = S.n, tid = threadIdx.x;
  for (size_t e = tid; e < 3 * S.vsize(); e += blockDim.x) S.Xc[e] = 0.0;  // Xc, FG, GX are contiguous
  for (size_t e = tid; e < (size_t)kCh * S.ns; e += blockDim.x) S.W[e] = 0.0;
  __syncthreads();
  const int per = n * D;
  for (int e = tid; e < nr * per; e += blockDim.x) {
    const int r = e / per, rem = e - r * per;
    const int j = rem / D, d = rem - j * D;
    S.Xc[S.vix(d, r, j)] = X[(chain0 + r) * per + rem] - S.mu[d];
  }
  if (tid < kCh * D) {
    const int r = tid / D, d = tid - r * D;
    S.s[d * kCh + r] = r < nr ? sig_pre[(chain0 + r) * D + d] : 0.0;
  }
  if (tid >= 128 && tid < 128 + kCh * P) {
    const int t = tid - 128, r = t / P, k = t - r * P;
    S.tau[k * kCh + r] = r < nr ? th_pre[(chain0 + r) * P + k] : 0.0;
  }
}
