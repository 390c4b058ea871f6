// Fast path of the posterior core for the headline shapes: np = 8*ceil(n/8) <= 168 (n <= 168 grid
// points, e.g. the reference's SEIR setting n = 161) and D <= 4 components.
//
// Same math and the same DMMA contractions as posterior_core.cuh, but laid out so that shared memory
// holds only what the tensor-core B operands need.  Measured on B200: the number of matrix bytes a CTA
// can keep in flight is bounded by the L1 capacity left over by the shared-memory carve-out
// (profiles/r01_notes.md), so every array that is not a B operand moves out of shared memory:
//   * one warp owns one 8-row block of every matrix; lane = 4g + c owns grid index j = 8*warp + g of
//     chains 2c and 2c+1 ("own elements") in ALL phases, pointwise ones included;
//   * the gradient dL/dX lives in registers (2*D doubles per lane) from the first contraction to the
//     final store / momentum kick;
//   * momentum and the saved start point of an HMC transition live in a per-CTA global scratch slot
//     in own-element order (coalesced 16-byte accesses, L2-resident because the grid is persistent);
//   * shared memory: Xc, FG [D][8][ns], Wa, Wb [8][ns], y/mask, per-chain scalars  (~124 KB at n = 161).
// The grid is persistent (one CTA per SM looping over (dataset, chain-group) items); the first matrix
// fragments of the next evaluation -- of the next item, too -- are already in flight while the
// pointwise epilogue of the current one runs.
#pragma once
#include "posterior_core.cuh"

constexpr int kFastMaxNp = 168;
constexpr int kFastMaxD = 4;

// NP > 0: the padded grid size is a compile-time constant (all shared-memory offsets and trip counts fold
// into immediates -- the instantiation the headline n = 161 runs); NP = 0: run-time np <= kFastMaxNp.
template <class M, int NP>
struct FastScratch {
  static constexpr int NRED = 2 + M::D + M::P;            // t1, t2, SSE_d, sum_j vth_k
  static constexpr int kCtl = 16 + 2 * M::P + 2 * M::D;   // as Scratch<M>::kCtl
  static constexpr int kSmall = kCh * (5 * M::P + 5 * M::D + 1 + kCtl + NRED) + 3 * M::D;
  static_assert(NRED <= 16, "per-warp partial sums alias Wa|Wb: needs NRED <= 16");
  double* base;
  int n, np_rt;
  __device__ __forceinline__ int np() const { return NP > 0 ? NP : np_rt; }
  __device__ __forceinline__ int ns() const { return magi_chain_stride(np()); }
  __device__ __forceinline__ int nblk() const { return np() >> 3; }
  __device__ __forceinline__ size_t vsz() const { return (size_t)M::D * kCh * ns(); }
  __device__ __forceinline__ double* Xc() const { return base; }                       // [D][8][ns]
  __device__ __forceinline__ double* FG() const { return base + vsz(); }               // [D][8][ns]
  __device__ __forceinline__ double* Wa() const { return base + 2 * vsz(); }           // [8][ns]
  __device__ __forceinline__ double* Wb() const { return Wa() + (size_t)kCh * ns(); }  // [8][ns]
  __device__ __forceinline__ double* wpart() const { return Wa(); }  // [nblk][8][NRED], aliases Wa|Wb (dead then)
  __device__ __forceinline__ double* Y() const { return Wb() + (size_t)kCh * ns(); }   // [D][np]
  __device__ __forceinline__ double* MK() const { return Y() + M::D * np(); }          // [D][np]
  __device__ __forceinline__ double* sm() const { return MK() + M::D * np(); }
  __device__ __forceinline__ double* tau() const { return sm(); }                      // [P][8] x5
  __device__ __forceinline__ double* th() const { return sm() + 1 * M::P * kCh; }
  __device__ __forceinline__ double* sgt() const { return sm() + 2 * M::P * kCh; }
  __device__ __forceinline__ double* ptau() const { return sm() + 3 * M::P * kCh; }
  __device__ __forceinline__ double* gtau() const { return sm() + 4 * M::P * kCh; }
  __device__ __forceinline__ double* s() const { return sm() + 5 * M::P * kCh; }       // [D][8] x5
  __device__ __forceinline__ double* sig2() const { return s() + 1 * M::D * kCh; }
  __device__ __forceinline__ double* sgs() const { return s() + 2 * M::D * kCh; }
  __device__ __forceinline__ double* ps() const { return s() + 3 * M::D * kCh; }
  __device__ __forceinline__ double* gs() const { return s() + 4 * M::D * kCh; }
  __device__ __forceinline__ double* L() const { return s() + 5 * M::D * kCh; }        // [8]
  __device__ __forceinline__ double* mu() const { return L() + kCh; }                  // [D] x3
  __device__ __forceinline__ double* Nd() const { return mu() + M::D; }
  __device__ __forceinline__ double* LB() const { return mu() + 2 * M::D; }
  __device__ __forceinline__ double* tot() const { return mu() + 3 * M::D; }           // [8][NRED]
  __device__ __forceinline__ double* ctl() const { return tot() + kCh * NRED; }        // [kCtl][8]

  __host__ __device__ static size_t elems(int np_) {
    const int ns_ = magi_chain_stride(np_);
    const size_t e = (size_t)2 * M::D * kCh * ns_ + (size_t)2 * kCh * ns_ + (size_t)2 * M::D * np_ + kSmall;
    return (e + 1) & ~(size_t)1;
  }
  __device__ __forceinline__ size_t vix(int d, int r, int j) const { return ((size_t)d * kCh + r) * ns() + j; }
};

// per-CTA global scratch slot: arrays in own-element order  [D][warp][lane][2]
template <class M>
__host__ __device__ inline size_t fast_slot_elems(int np) { return (size_t)M::D * (np >> 3) * 64; }

__device__ __forceinline__ size_t own_ix(int d, int nblk) {  // index of this lane's pair for component d
  return ((size_t)(d * nblk + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31)) * 2;
}

// fold a value over the 8 lanes sharing c = lane & 3 (i.e. over g): lanes 0..3 end up with the sum
__device__ __forceinline__ double fold_g(double v) {
  v += magi_shfl_xor(v, 4);
  v += magi_shfl_xor(v, 8);
  v += magi_shfl_xor(v, 16);
  return v;
}

// Evaluate base log-posterior L and gradient at the state in shared memory (S.Xc(), S.tau(), S.s()).
// On entry a[] holds this warp's first kU steps of sym(C^-1) of component 0.  On exit S.L(), S.gs(),
// S.gtau() are set, gxr[d][q] = dL/dX[j, d] of chain 2c+q at the lane's own grid index (garbage for
// j >= n), and a[] holds the first steps of `next_mats` (if not null).  All threads must call.
template <class M, int NP>
__device__ void fast_eval(const FastScratch<M, NP>& S, const double* __restrict__ mats,
                          const double* __restrict__ next_mats, double inv_beta, double (&gxr)[M::D][2],
                          double2 (&a)[kU]) {
  constexpr int D = M::D, P = M::P, NRED = FastScratch<M, NP>::NRED;
  const int n = S.n, np = S.np(), ns = S.ns();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = S.nblk();
  const int g = lane >> 2, c2 = 2 * (lane & 3);
  const int j = warp * 8 + g;
  const size_t msz = (size_t)np * np;
  const int nblk = S.nblk();
  const StepRange rg{0, nblk};  // all column steps: skipping the all-zero tiles of banded matrices measured
                                // slower here (unbalanced warps, run-time trip counts; profiles/r01_notes.md)
  const bool valid = j < n;
  const size_t o0 = (size_t)c2 * ns + j, o1 = o0 + ns;  // own elements inside one [8][ns] array

  // per-chain transforms (threads 0..7)
  if (tid < kCh) {
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double t = S.tau()[k * kCh + tid];
      S.th()[k * kCh + tid] = magi_softplus(t);
      S.sgt()[k * kCh + tid] = magi_sigmoid(t);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double z = S.s()[d * kCh + tid];
      S.sig2()[d * kCh + tid] = magi_softplus(z) + S.LB()[d];
      S.sgs()[d * kCh + tid] = magi_sigmoid(z);
    }
  }
  __syncthreads();

  // f(X, theta) at the own grid index of both chains
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    double th[P], x[D], f[D];
#pragma unroll
    for (int k = 0; k < P; ++k) th[k] = S.th()[k * kCh + c2 + q];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = S.Xc()[S.vix(d, c2 + q, j)] + S.mu()[d];
    M::f(x, th, f);
#pragma unroll
    for (int d = 0; d < D; ++d) S.FG()[S.vix(d, c2 + q, j)] = valid ? f[d] : 0.0;
  }
  __syncthreads();

  double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
  // A(d): u = S_C xc (-> gxr = 2u, t1 += xc.u) ; w = m xc -> Wa
  auto pass_a = [&](int d) {
    const double* SC = mats + (size_t)(3 * d + 0) * msz;
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    const double* xc = S.Xc() + S.vix(d, 0, 0);
    double c0, c1;
    const double* sm = stream_ptr<kFwd>(Mm, np, warp, lane);
    mma_task<false, kFwd, kFwd>(a, stream_ptr<kFwd>(SC, np, warp, lane), rg, sm, rg, xc, nullptr, ns, np, c0, c1);
    gxr[d][0] = 2.0 * c0;
    gxr[d][1] = 2.0 * c1;
    t1[0] = fma(xc[o0], c0, t1[0]);
    t1[1] = fma(xc[o1], c1, t1[1]);
    mma_task<false, kFwd, kFwd>(a, sm, rg, stream_ptr<kFwd>(SK, np, warp, lane), rg, xc, nullptr, ns, np, c0, c1);
    S.Wa()[o0] = c0;
    S.Wa()[o1] = c1;
  };
  pass_a(0);
  __syncthreads();
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    double* fg = S.FG() + S.vix(d, 0, 0);
    double g0, g1;
    {  // B(d): g = 2 S_K (f - w) -> Wb ; t2 += (f - w) . S_K (f - w)
      double c0, c1;
      mma_task<true, kFwd, kTr>(a, stream_ptr<kFwd>(SK, np, warp, lane), rg, stream_ptr<kTr>(Mm, np, warp, lane), rg,
                                fg, S.Wa(), ns, np, c0, c1);
      g0 = 2.0 * c0;
      g1 = 2.0 * c1;
      S.Wb()[o0] = g0;
      S.Wb()[o1] = g1;
      t2[0] = fma(fg[o0] - S.Wa()[o0], c0, t2[0]);
      t2[1] = fma(fg[o1] - S.Wa()[o1], c1, t2[1]);
    }
    __syncthreads();
    {  // C(d): gxr -= m^T g (second read of m: L2) ; FG_d <- g ; A(d+1)
      double c0, c1;
      const bool last = d + 1 == D;
      const double* nm = last ? next_mats : mats + (size_t)(3 * (d + 1)) * msz;
      mma_task<false, kTr, kFwd>(a, stream_ptr<kTr>(Mm, np, warp, lane), rg,
                                 nm ? stream_ptr<kFwd>(nm, np, warp, lane) : nullptr, rg, S.Wb(), nullptr, ns, np,
                                 c0, c1);
      gxr[d][0] -= c0;
      gxr[d][1] -= c1;
      fg[o0] = g0;
      fg[o1] = g1;
      if (!last) pass_a(d + 1);
    }
    __syncthreads();
  }

  // pointwise epilogue at the own elements: ODE Jacobian terms, likelihood, assemble the gradient
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int ch = c2 + q;
    double th[P], x[D], gg[D], vx[D], vth[P], red[NRED];
#pragma unroll
    for (int k = 0; k < P; ++k) th[k] = S.th()[k * kCh + ch];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      x[d] = S.Xc()[S.vix(d, ch, j)] + S.mu()[d];
      gg[d] = S.FG()[S.vix(d, ch, j)];   // zero for j >= n
    }
    M::vjp(x, th, gg, vx, vth);
    red[0] = t1[q];
    red[1] = t2[q];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double e = (valid && S.MK()[d * np + j] != 0.0) ? x[d] - S.Y()[d * np + j] : 0.0;
      red[2 + d] = e * e;
      gxr[d][q] = -0.5 * ((gxr[d][q] + vx[d]) * inv_beta + 2.0 * e / S.sig2()[d * kCh + ch]);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) red[2 + D + k] = valid ? vth[k] : 0.0;
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      const double v = fold_g(red[k]);
      if (lane < 4) S.wpart()[((size_t)warp * kCh + ch) * NRED + k] = v;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int e = tid, nthr = 32 * nblk; e < kCh * NRED; e += nthr) {  // (chain, k) totals over the warps in a fixed order
    double v = 0.0;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) v += S.wpart()[(size_t)w * kCh * NRED + e];
    S.tot()[e] = v;
  }
  __syncthreads();
  if (tid < kCh) {
    const int c = tid;
    const double* tot = S.tot() + c * NRED;
    double t3 = 0.0, t4 = 0.0, lj = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double s2 = S.sig2()[d * kCh + c], sg = S.sgs()[d * kCh + c], z = S.s()[d * kCh + c];
      t3 += S.Nd()[d] * log(2.0 * M_PI * s2);
      t4 += tot[2 + d] / s2;
      lj += z - magi_softplus(z);
      S.gs()[d * kCh + c] = -0.5 * (S.Nd()[d] / s2 - tot[2 + d] / (s2 * s2)) * sg + (1.0 - sg);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double t = S.tau()[k * kCh + c], sg = S.sgt()[k * kCh + c];
      lj += t - magi_softplus(t);
      S.gtau()[k * kCh + c] = -0.5 * inv_beta * tot[2 + D + k] * sg + (1.0 - sg);
    }
    S.L()[c] = -0.5 * ((inv_beta * (tot[0] + tot[1])) + (t3 + t4)) + lj;
  }
  __syncthreads();
}

// Per-item loads: dataset constants, then the chain states (reference layout X[n][D] per chain) into
// shared memory, centred; chains >= nr and grid indices >= n are zero.  Ends with __syncthreads().
template <class M, int NP>
__device__ void fast_load_item(const FastScratch<M, NP>& S, const magi_problem_t& pb, int b, const double* X,
                               const double* sig_pre, const double* th_pre, size_t chain0, int nr) {
  constexpr int D = M::D, P = M::P;
  const int n = S.n, np = S.np(), tid = threadIdx.x, nthr = 32 * S.nblk();
  if (tid < D) {
    S.mu()[tid] = pb.mu[(size_t)b * D + tid];
    S.Nd()[tid] = pb.N_ds[(size_t)b * D + tid];
    S.LB()[tid] = pb.LB[(size_t)b * D + tid];
  }
#pragma unroll 1
  for (int e = tid; e < D * np; e += nthr) {
    const int d = e / np, jj = e - d * np;
    double yv = 0.0, mk = 0.0;
    if (jj < n) {
      const size_t ai = ((size_t)b * n + jj) * D + d;
      mk = pb.mask[ai] ? 1.0 : 0.0;
      yv = mk != 0.0 ? pb.y[ai] : 0.0;
    }
    S.Y()[e] = yv;
    S.MK()[e] = mk;
  }
  // chain states at the own elements (lane 4g+c: grid index 8*warp + g of chains 2c, 2c+1): 32 contiguous
  // bytes per (chain, j) for D = 4.  Every (chain, j < np) entry of Xc is written -- zeros for chains >= nr
  // and for the padding j >= n (B-operand padding must be finite); FG, Wa, Wb are fully rewritten by every
  // evaluation, so nothing needs clearing.
  {
    const int lane = tid & 31, warp = tid >> 5, j = warp * 8 + (lane >> 2), c2 = 2 * (lane & 3);
    double mu[D];
#pragma unroll
    for (int d = 0; d < D; ++d) mu[d] = pb.mu[(size_t)b * D + d];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = c2 + q;
      const bool ok = ch < nr && j < n;
      const double* xp = X + ((chain0 + (ok ? ch : 0)) * n + (ok ? j : 0)) * D;
#pragma unroll
      for (int d = 0; d < D; ++d) S.Xc()[S.vix(d, ch, j)] = ok ? xp[d] - mu[d] : 0.0;
    }
  }
  if (tid < kCh * D) {
    const int r = tid / D, d = tid - r * D;
    S.s()[d * kCh + r] = r < nr ? sig_pre[(chain0 + r) * D + d] : 0.0;
  }
  if (tid >= 32 * (nthr > 32) && tid < 32 * (nthr > 32) + kCh * P) {
    const int t = tid - 32 * (nthr > 32), r = t / P, k = t - r * P;
    S.tau()[k * kCh + r] = r < nr ? th_pre[(chain0 + r) * P + k] : 0.0;
  }
  __syncthreads();
}
