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
#ifndef MAGI_KU
#define MAGI_KU 7
#endif
#ifndef MAGI_LDG_MODE
#define MAGI_LDG_MODE 0   // 0: ld.global.nc  1: ld.global.nc.L1::no_allocate  2: ld.global.cg
#endif
constexpr int kU = MAGI_KU;                 // 8-column steps per register batch (kU x 16 B per lane in flight)

// chain stride of the shared-memory vector arrays, ns = 8 (mod 16) doubles: the 128-bit B-fragment loads are
// served a quarter warp at a time (lanes 4g+c, g in {2q, 2q+1}: two chains x four 16-byte column pairs);
// with the odd chain 64 B (mod 128 B) away from the even one the 8 lanes cover all 32 banks exactly once.
// (ns = np + 2 was measured 2-way conflicted on every B load: ncu "L1 Wavefronts Shared Excessive".)
__host__ __device__ static inline int magi_chain_stride(int np) { return (np & 15) == 8 ? np : np + 8; }

// Scratch arrays of one CTA.  Vector arrays are [D][8][ns] (chain-major, grid index fastest).
template <class M>
struct Scratch {
  int n, np, ns;
  double *Xc, *FG, *GX;      // [D][8][ns]   centred state | f -> residual r -> g = 2 S_K r | gradient
  double* PX;                // [D][8][ns]   momentum (HMC kernels only; else nullptr)
  double *Wa0, *Wa1, *Wb;        // [8][ns] x3   m_d xc_d (double-buffered over d) and g_d = 2 S_K r_d
  double *Y, *MK;            // [D][np]      observations / mask as 0.0 or 1.0
  double *tau, *th, *sgt;    // [P][8]       theta_pre, softplus, sigmoid
  double *s, *sig2, *sgs;    // [D][8]       sigma_pre, softplus + LB, sigmoid
  double *ptau, *ps;         // [P][8], [D][8] momenta of the small state parts
  double *gtau, *gs;         // [P][8], [D][8] base gradient w.r.t. tau, s
  double* L;                 // [8]          base log-posterior
  double *mu, *Nd, *LB;      // [D]
  double* wpart;             // [kMaxWarps][NRED]
  double* wt1;               // [kMaxWarps][8]  per-warp partial sums of t1
  double* ctl;               // [kCtl][8]    sampler control values (step size, energies, saved small state)
  static constexpr int kCtl = 16 + 2 * M::P + 2 * M::D;
  static constexpr int NRED = 2 + M::D + M::P;

  __host__ __device__ static size_t big_elems(int np, bool with_momentum) {
    return (size_t)(with_momentum ? 4 : 3) * M::D * kCh * magi_chain_stride(np) + (size_t)3 * kCh * magi_chain_stride(np);
  }
  __host__ __device__ static size_t small_elems(int np) {
    const size_t e = (size_t)2 * M::D * np + (size_t)kCh * (5 * M::P + 5 * M::D + 1 + kCtl) + 3 * M::D +
                     (size_t)kMaxWarps * (NRED + kCh);
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
    Wa0 = p; Wa1 = p + (size_t)kCh * ns; Wb = p + (size_t)2 * kCh * ns;
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
    wt1 = p; p += kMaxWarps * kCh;
    ctl = p;
  }
  // Tell the compiler that every scratch pointer is a shared-memory address (LDS/STS instead of
  // generic loads).  Only valid when the big arrays were carved out of shared memory too.
  __device__ __forceinline__ void assume_shared() const {
#define MAGI_ASSUME_SHARED(q) __builtin_assume(__isShared(q))
    MAGI_ASSUME_SHARED(Xc); MAGI_ASSUME_SHARED(FG); MAGI_ASSUME_SHARED(GX); MAGI_ASSUME_SHARED(Wa0);
    MAGI_ASSUME_SHARED(Wa1); MAGI_ASSUME_SHARED(Wb); MAGI_ASSUME_SHARED(Y); MAGI_ASSUME_SHARED(MK);
    MAGI_ASSUME_SHARED(tau); MAGI_ASSUME_SHARED(th); MAGI_ASSUME_SHARED(sgt); MAGI_ASSUME_SHARED(s);
    MAGI_ASSUME_SHARED(sig2); MAGI_ASSUME_SHARED(sgs); MAGI_ASSUME_SHARED(ptau); MAGI_ASSUME_SHARED(ps);
    MAGI_ASSUME_SHARED(gtau); MAGI_ASSUME_SHARED(gs); MAGI_ASSUME_SHARED(L); MAGI_ASSUME_SHARED(mu);
    MAGI_ASSUME_SHARED(Nd); MAGI_ASSUME_SHARED(LB); MAGI_ASSUME_SHARED(wpart); MAGI_ASSUME_SHARED(wt1);
    MAGI_ASSUME_SHARED(ctl);
    if (PX) MAGI_ASSUME_SHARED(PX);
#undef MAGI_ASSUME_SHARED
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

// A matrix stream as seen by one lane.  Forward (TR = 0; y = A x, 8 output rows i0..i0+7): the lane walks
// row i0 + g in 8-column steps, 16 B at columns 8s + 2c, 2c+1.  Transposed (TR = 1; y = A^T x, 8 output
// columns j0..j0+7): column j0 + g, rows 8s + 2c and 8s + 2c + 1.  Either way step s contributes two
// k=4 contractions whose k-index c maps to vector element 8s + 2c (+1), i.e. the B fragments are the
// 16 B at x[chain g][8s + 2c].
constexpr int kFwd = 0, kTr = 1, kNone = -1;

// Packed matrices are stored TILED: [row block rb][column step s] tiles of 8 x 8 doubles (512 B, row-major
// inside the tile), so that one warp-wide 128-bit load reads one whole contiguous tile (lane l <- bytes
// 16 l .. 16 l + 15, i.e. T[g][2c], T[g][2c+1]) and a warp streams 8 rows of a matrix as one contiguous run
// of np/8 tiles.  A transposed stream reads the tiles of one block COLUMN (stride np/8 tiles) with the
// same coalesced load and transposes the fragment in registers (transpose_frag) when it is consumed.
template <int TR>
__device__ __forceinline__ const double* stream_ptr(const double* A, int np, int blk, int lane) {
  const int lo = 2 * magi_tile_slot(lane >> 2, lane & 3);   // this lane's pair (g, c) inside a tile (common.cuh)
  return TR == kTr ? A + (size_t)blk * 64 + lo : A + (size_t)blk * (np >> 3) * 64 + lo;
}

// Lane 4g+c holds (T[g][2c], T[g][2c+1]) of an 8x8 tile; returns (T[2c][g], T[2c+1][g]) -- the A fragments
// of the two k=4 contractions of a TRANSPOSED product.  Three 64-bit shuffles: a 2x2 exchange between the
// lanes of rows 2r / 2r+1, then one gather.
__device__ __forceinline__ double2 transpose_frag(double2 t) {
  const int lane = threadIdx.x & 31;
  const bool odd = (lane >> 2) & 1;
  const double recv = magi_shfl_xor(odd ? t.x : t.y, 4);
  const double2 w = odd ? make_double2(recv, t.y) : make_double2(t.x, recv);  // even row: (x, x') ; odd row: (y, y')
  const int src = ((2 * (lane & 3) + ((lane >> 2) & 1)) << 2) | (lane >> 3);
  return make_double2(__shfl_sync(MAGI_FULL_MASK, w.x, src), __shfl_sync(MAGI_FULL_MASK, w.y, src));
}

__device__ __forceinline__ double2 ldg_f64x2(const double* p) {
#if MAGI_LDG_MODE == 1
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#elif MAGI_LDG_MODE == 2
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const double2*>(p));
#endif
}
__device__ __forceinline__ double ldg_f64(const double* p) {
#if MAGI_LDG_MODE == 1
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#elif MAGI_LDG_MODE == 2
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

// Column steps (forward) / row steps (transposed) of a block that can hold non-zeros when the matrices
// are banded (tf.linalg.band_part, magi_v2.py:271-274): tile (blk, s) is non-zero only if
// |s - blk| <= kb with kb = floor((band + 7) / 8).  Dense matrices: kb >= number of blocks.
struct StepRange {
  int lo, hi;  // steps lo .. hi-1
};
__device__ __forceinline__ int band_blocks(int band, int nblk) { return band < 0 ? nblk : (band + 7) >> 3; }
__device__ __forceinline__ StepRange band_range(int blk, int nblk, int kb) {
  return StepRange{max(0, blk - kb), min(nblk, blk + kb + 1)};
}

// step s of a stream (zero at and beyond step `hi`)
template <int TR>
__device__ __forceinline__ double2 load_step(const double* p, int s, int hi, int np) {
  double2 v = make_double2(0.0, 0.0);
  if (s < hi) {
    v = ldg_f64x2(TR == kTr ? p + (size_t)s * (np >> 3) * 64 : p + 64 * s);
  }
  return v;
}

template <int TR, int U>
__device__ __forceinline__ void load_batch(double2 (&a)[U], const double* p, StepRange r, int np) {
#pragma unroll
  for (int u = 0; u < U; ++u) a[u] = load_step<TR>(p, r.lo + u, r.hi, np);
}

// One 8-row (CT = kFwd) or 8-column (CT = kTr) block of a contraction with the 8 chain vectors x[8][ns]
// (or x - xsub when SUB) over the steps `cr`: a rolling register pipeline.  On entry a[u] holds step
// cr.lo + u of the current stream `cp` (u < U); every register is refilled with step s + kU as soon as
// step s has been consumed, so U 16-byte loads per lane stay in flight; during the last batch the
// refills switch to the warp's NEXT task (`nxp`, kind NT, steps `nr`; possibly in a later phase, behind
// a __syncthreads; nullptr: no next task), whose first U steps are therefore already in flight when
// it starts.  Result: c0, c1 = y[chain 2c], y[chain 2c+1] at block element g (lane = 4g + c).
template <bool SUB, int CT, int NT, int U>
__device__ __forceinline__ void mma_task(double2 (&a)[U], const double* cp, StepRange cr, const double* nxp,
                                         StepRange nr, const double* x, const double* xsub, int ns, int np,
                                         double& c0, double& c1) {
  const int lane = threadIdx.x & 31;
  const size_t bo = (size_t)(lane >> 2) * ns + 2 * (lane & 3);
  const double* bp = x + bo;
  const double* bs = SUB ? xsub + bo : nullptr;
  const int nsteps = cr.hi - cr.lo;
  const int nlast = cr.lo + ((nsteps - 1) / U) * U;  // first step of the last batch
  double acc[2][2];  // two independent accumulator pairs (even / odd k-group of a step)
#pragma unroll
  for (int q = 0; q < 2; ++q) acc[q][0] = acc[q][1] = 0.0;
  auto consume = [&](int s, int u) {
    double2 b = *reinterpret_cast<const double2*>(bp + 8 * s);
    if (SUB) {
      const double2 b2 = *reinterpret_cast<const double2*>(bs + 8 * s);
      b.x -= b2.x;
      b.y -= b2.y;
    }
    const double2 av = CT == kTr ? transpose_frag(a[u]) : a[u];
    dmma(acc[0][0], acc[0][1], av.x, b.x);
    dmma(acc[1][0], acc[1][1], av.y, b.y);
  };
  // all batches but the last: refill from the current stream.  (Two copies of the body on purpose: with
  // compile-time ranges every predicate folds away, which measured ~20 % faster than one shared body.)
#pragma unroll 1
  for (int s0 = cr.lo; s0 < nlast; s0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      consume(s0 + u, u);
      a[u] = load_step<CT>(cp, s0 + u + U, cr.hi, np);
    }
  }
  const int nhi = nxp != nullptr ? nr.hi : 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (nlast + u < cr.hi) consume(nlast + u, u);
    a[u] = load_step<NT>(nxp, nr.lo + u, nhi, np);
  }
  c0 = acc[0][0] + acc[1][0];
  c1 = acc[0][1] + acc[1][1];
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
__device__ void eval_logpost_grad(const Scratch<M>& S, const double* __restrict__ mats, double inv_beta,
                                  int band) {
  constexpr int D = M::D, P = M::P, NRED = Scratch<M>::NRED;
  const int n = S.n, np = S.np, ns = S.ns;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const ChainMap cm = chain_map();
  const int r = cm.r;
  const size_t msz = (size_t)np * np;
  const int nblk = np >> 3;  // 8-row blocks = 8-column steps
  const int g = lane >> 2, c2 = 2 * (lane & 3);
  const int kb = band_blocks(band, nblk);
  auto rng = [&](int blk) { return band_range(blk, nblk, kb); };

  chain_scalars(S);
  // first matrix batch of this warp goes in flight before anything else
  double2 a[kU];
  const bool active = warp < nblk;
  if (active) load_batch<kFwd>(a, stream_ptr<kFwd>(mats, np, warp, lane), rng(warp), np);
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

  // Phases per component d (two barriers each):
  //   A(d): u = 2 S_C xc -> GX_d ; w = m xc -> Wa[d&1]                      (A(0) alone, else merged into C(d-1))
  //   B(d): g = 2 S_K (f - w) -> Wb        (the residual r = f - w is formed on the fly as the B operand)
  //   C(d): GX_d -= m^T g (second read of m: L2) | t1, t2 partial sums, FG_d <- g | A(d+1)
  double t1a = 0.0, t1b = 0.0, t2 = 0.0;
  auto pass_a = [&](int d, bool chain_next) {
    const double* SC = mats + (size_t)(3 * d + 0) * msz;
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    const double* xc = S.Xc + S.vix(d, 0, 0);
    double* gx = S.GX + S.vix(d, 0, 0);
    double* wa = (d & 1) ? S.Wa1 : S.Wa0;
    for (int blk = warp; blk < nblk; blk += nw) {
      double c0, c1;
      const double* sm = stream_ptr<kFwd>(Mm, np, blk, lane);
      mma_task<false, kFwd, kFwd>(a, stream_ptr<kFwd>(SC, np, blk, lane), rng(blk), sm, rng(blk), xc, nullptr, ns, np,
                                  c0, c1);
      gx[(size_t)c2 * ns + blk * 8 + g] = 2.0 * c0;
      gx[(size_t)(c2 + 1) * ns + blk * 8 + g] = 2.0 * c1;
      t1a = fma(xc[(size_t)c2 * ns + blk * 8 + g], c0, t1a);  // xc . S_C xc for chains 2c, 2c+1
      t1b = fma(xc[(size_t)(c2 + 1) * ns + blk * 8 + g], c1, t1b);
      const bool more = blk + nw < nblk;
      const double* nx = more ? stream_ptr<kFwd>(SC, np, blk + nw, lane)
                              : (chain_next ? stream_ptr<kFwd>(SK, np, warp, lane) : nullptr);
      mma_task<false, kFwd, kFwd>(a, sm, rng(blk), nx, rng(more ? blk + nw : warp), xc, nullptr, ns, np, c0, c1);
      wa[(size_t)c2 * ns + blk * 8 + g] = c0;
      wa[(size_t)(c2 + 1) * ns + blk * 8 + g] = c1;
    }
  };
  pass_a(0, true);
  __syncthreads();
  for (int d = 0; d < D; ++d) {
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    double* fg = S.FG + S.vix(d, 0, 0);
    double* gx = S.GX + S.vix(d, 0, 0);
    const double* wa = (d & 1) ? S.Wa1 : S.Wa0;
    // B(d)
    for (int blk = warp; blk < nblk; blk += nw) {
      double c0, c1;
      const double* cur = stream_ptr<kFwd>(SK, np, blk, lane);
      if (blk + nw < nblk)
        mma_task<true, kFwd, kFwd>(a, cur, rng(blk), stream_ptr<kFwd>(SK, np, blk + nw, lane), rng(blk + nw), fg, wa,
                                   ns, np, c0, c1);
      else
        mma_task<true, kFwd, kTr>(a, cur, rng(blk), stream_ptr<kTr>(Mm, np, warp, lane), rng(warp), fg, wa, ns, np,
                                  c0, c1);
      S.Wb[(size_t)c2 * ns + blk * 8 + g] = 2.0 * c0;
      S.Wb[(size_t)(c2 + 1) * ns + blk * 8 + g] = 2.0 * c1;
    }
    __syncthreads();
    // C(d)
    for (int blk = warp; blk < nblk; blk += nw) {
      double c0, c1;
      const double* cur = stream_ptr<kTr>(Mm, np, blk, lane);
      if (blk + nw < nblk)
        mma_task<false, kTr, kTr>(a, cur, rng(blk), stream_ptr<kTr>(Mm, np, blk + nw, lane), rng(blk + nw), S.Wb,
                                  nullptr, ns, np, c0, c1);
      else
        mma_task<false, kTr, kFwd>(a, cur, rng(blk),
                                   d + 1 < D ? stream_ptr<kFwd>(mats + (size_t)(3 * (d + 1)) * msz, np, warp, lane)
                                             : nullptr,
                                   rng(warp), S.Wb, nullptr, ns, np, c0, c1);
      gx[(size_t)c2 * ns + blk * 8 + g] -= c0;
      gx[(size_t)(c2 + 1) * ns + blk * 8 + g] -= c1;
    }
    if (d + 1 < D) pass_a(d + 1, true);
    for (int j = cm.j0; j < n; j += cm.jstride) {
      const size_t e = (size_t)r * ns + j;
      const double gv = S.Wb[e];
      t2 = fma(fg[e] - wa[e], 0.5 * gv, t2);
      fg[e] = gv;
    }
    __syncthreads();
  }

  // pointwise: ODE Jacobian terms, likelihood, assemble gradient
  double red[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) red[k] = 0.0;
  red[1] = t2;
  // t1: lanes hold partial sums for chains 2c, 2c+1 over their block rows g; fold the 8 values of g
  t1a += magi_shfl_xor(t1a, 4); t1a += magi_shfl_xor(t1a, 8); t1a += magi_shfl_xor(t1a, 16);
  t1b += magi_shfl_xor(t1b, 4); t1b += magi_shfl_xor(t1b, 8); t1b += magi_shfl_xor(t1b, 16);
  if (lane < 4) {
    S.wt1[warp * kCh + 2 * lane] = t1a;
    S.wt1[warp * kCh + 2 * lane + 1] = t1b;
  }
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
    {
      double v = 0.0;
      for (int w = 0; w < nw; ++w) v += S.wt1[w * kCh + c];
      tot[0] = v;
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
  for (size_t e = tid; e < (size_t)3 * kCh * S.ns; e += blockDim.x) S.Wa0[e] = 0.0;
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
