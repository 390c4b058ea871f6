// Fast path of the posterior core for the headline shapes: np = 8*ceil(n/8) <= 168 (n <= 168 grid
// points, e.g. the reference's SEIR setting n = 161) and D <= 4 components.
//
// Same math and the same DMMA contractions as posterior_core.cuh, organised around what was measured on
// B200 (profiles/r01_notes.md): a CTA saturates its share of the L2->SM fabric with ~40 KB of 128-bit
// loads in flight (tools/stream_probe.cu), so few warps with deep register pipelines beat many warps;
// what costs time is everything that is not streaming -- index arithmetic, spills, barriers, per-item
// prologues -- and every byte that crosses the fabric, L2 hits included.  Hence:
//   * nw = ceil(nblk / 2) warps (11 at n = 161, 168 registers each); warp w owns the 8-row blocks w and
//     w + nw of every matrix, a pairing that also balances the band (magi_v2.py:271-274): with bandsize
//     80 of n = 161 every warp streams 31 of its 42 tiles per matrix pass, the others are all-zero and
//     never read;
//   * lane = 4g + c owns grid indices j_h = 8*block_h + g of chains 2c and 2c+1 ("own elements") in ALL
//     phases, pointwise ones included; the gradient dL/dX lives in registers from the first contraction
//     to the final store / momentum kick; momentum and the saved start point of an HMC transition live
//     in a per-CTA global scratch slot in own-element order (coalesced, L2-resident: persistent grid);
//   * shared memory holds only the tensor-core B operands (Xc, FG [D][8][ns], Wa, Wb [8][ns]) and the
//     per-chain scalars; y / mask are read from global memory at the own elements;
//   * NP > 0: the padded grid size is a compile-time constant, all shared-memory offsets fold into
//     immediates (the instantiation n in 161..168 runs); NP = 0: run-time np.
// The grid is persistent (one CTA per SM looping over (dataset, chain-group) items); the first matrix
// fragments of the next evaluation -- of the next item, too -- are already in flight while the
// pointwise epilogue of the current one runs.
#pragma once
#include "posterior_core.cuh"

// experiment knobs (tools/build_variant.sh only; never defined in the product build)
#ifdef MAGI_EXP_NOSYNC
#define MAGI_PHASE_SYNC() ((void)0)
#else
#define MAGI_PHASE_SYNC() __syncthreads()
#endif
#ifndef MAGI_FAST_KU
#define MAGI_FAST_KU 10
#endif

constexpr int kFastMaxNp = 168;
constexpr int kFastMaxD = 4;
constexpr int kFU = MAGI_FAST_KU;                       // pipeline depth of the fast path (16-byte loads per lane)
constexpr int kFastMaxWarps = (kFastMaxNp / 8 + 1) / 2; // 11
constexpr int kFastMaxThreads = 32 * kFastMaxWarps;     // 352

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
  __device__ __forceinline__ int nw() const { return (nblk() + 1) >> 1; }
  __device__ __forceinline__ size_t vsz() const { return (size_t)M::D * kCh * ns(); }
  __device__ __forceinline__ double* Xc() const { return base; }                       // [D][8][ns]
  __device__ __forceinline__ double* FG() const { return base + vsz(); }               // [D][8][ns]
  __device__ __forceinline__ double* Wa() const { return base + 2 * vsz(); }           // [8][ns]
  __device__ __forceinline__ double* Wb() const { return Wa() + (size_t)kCh * ns(); }  // [8][ns]
  __device__ __forceinline__ double* wpart() const { return Wa(); }  // [nw][8][NRED], aliases Wa|Wb (dead then)
  __device__ __forceinline__ double* sm() const { return Wb() + (size_t)kCh * ns(); }
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
    const size_t e = (size_t)2 * M::D * kCh * ns_ + (size_t)2 * kCh * ns_ + kSmall;
    return (e + 1) & ~(size_t)1;
  }
  __device__ __forceinline__ size_t vix(int d, int r, int j) const { return ((size_t)d * kCh + r) * ns() + j; }
};

// per-CTA global scratch slot: arrays in own-element order  [D][h][warp][lane][2]
template <class M>
__host__ __device__ inline size_t fast_slot_elems(int np) { return (size_t)M::D * 2 * (((np >> 3) + 1) >> 1) * 64; }

__device__ __forceinline__ size_t own_ix(int d, int h, int nw) {  // this lane's chain pair for (component d, block h)
  return ((size_t)((d * 2 + h) * nw + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31)) * 2;
}

// fold a value over the 8 lanes sharing c = lane & 3 (i.e. over g): lanes 0..3 end up with the sum
__device__ __forceinline__ double fold_g(double v) {
  v += magi_shfl_xor(v, 4);
  v += magi_shfl_xor(v, 8);
  v += magi_shfl_xor(v, 16);
  return v;
}

// What a lane owns: its two row blocks / grid indices.
struct Own {
  int b[2];   // row blocks (b[1] < 0: none)
  int j[2];   // grid indices 8*b[h] + g  (np for the missing block: never < n)
  int c2;     // first of its two chains
};
template <class M, int NP>
__device__ __forceinline__ Own fast_own(const FastScratch<M, NP>& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Own o;
  o.b[0] = warp;
  o.b[1] = warp + S.nw() < S.nblk() ? warp + S.nw() : -1;
  o.j[0] = 8 * o.b[0] + (lane >> 2);
  o.j[1] = o.b[1] >= 0 ? 8 * o.b[1] + (lane >> 2) : S.np();
  o.c2 = 2 * (lane & 3);
  return o;
}

// Evaluate base log-posterior L and gradient at the state in shared memory (S.Xc, S.tau, S.s).
// On entry a[] holds the warp's first kFU steps of sym(C^-1) of component 0, block b[0].  On exit
// S.L(), S.gs(), S.gtau() are set, gxr[h][d][q] = dL/dX[j_h, d] of chain 2c+q (garbage for j_h >= n),
// and a[] holds the first steps of `next_mats` (if not null).  yb / mb: y and mask of this dataset
// ([n][D], global).  All threads must call.
template <class M, int NP>
__device__ void fast_eval(const FastScratch<M, NP>& S, const double* __restrict__ mats,
                          const double* __restrict__ next_mats, double inv_beta, int band,
                          const double* __restrict__ yb, const uint8_t* __restrict__ mb,
                          double (&gxr)[2][M::D][2], double2 (&a)[kFU]) {
  constexpr int D = M::D, P = M::P, NRED = FastScratch<M, NP>::NRED;
  const int n = S.n, np = S.np(), ns = S.ns(), nblk = S.nblk(), nw = S.nw();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Own own = fast_own(S);
  const int c2 = own.c2;
  const bool has1 = own.b[1] >= 0;
  const size_t msz = (size_t)np * np;
  const int kb = band_blocks(band, nblk);
  const StepRange r0 = band_range(own.b[0], nblk, kb);
  const StepRange r1 = has1 ? band_range(own.b[1], nblk, kb) : StepRange{0, 0};
  // own elements inside one [8][ns] array: o[h][q]  (h = 1 aliases h = 0 when the lane has one block)
  size_t o[2][2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    o[h][0] = (size_t)c2 * ns + (has1 ? own.j[h] : own.j[0]);
    o[h][1] = o[h][0] + ns;
  }
  auto fwd = [&](const double* A, int h) { return stream_ptr<kFwd>(A, np, own.b[h], lane); };
  auto trp = [&](const double* A, int h) { return stream_ptr<kTr>(A, np, own.b[h], lane); };

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

  // f(X, theta) at the own grid indices of both chains
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h == 0 || has1) {
      const int j = own.j[h];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double th[P], x[D], f[D];
#pragma unroll
        for (int k = 0; k < P; ++k) th[k] = S.th()[k * kCh + c2 + q];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = S.Xc()[S.vix(d, c2 + q, j)] + S.mu()[d];
        M::f(x, th, f);
#pragma unroll
        for (int d = 0; d < D; ++d) S.FG()[S.vix(d, c2 + q, j)] = j < n ? f[d] : 0.0;
      }
    }
  }
  __syncthreads();

  double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
  // Per component d, three passes over its matrices (two barriers):
  //   A(d): u = S_C xc (-> g_x = 2u, t1 += xc.u) ; w = m xc -> Wa
  //   B(d): g = 2 S_K (f - w) -> Wb ; t2 += (f - w).S_K (f - w)     (r = f - w formed on the fly as B operand)
  //   C(d): g_x -= m^T g  (second read of m: an L2 hit) ; FG_d <- g
  // C(d-1) and A(d) share a barrier interval.  Every task leaves the first kFU steps of the warp's next
  // task in flight (across the barriers too).  A warp without a second block runs empty tasks for it.
  // The gradient of the component in flight is in `cur`; finished ones are pushed through gxr as a shift
  // register so that d stays a run-time loop variable (one copy of the code).
  double cur[2][2], gv[2][2];
#pragma unroll
  for (int h = 0; h < 2; ++h) cur[h][0] = cur[h][1] = gv[h][0] = gv[h][1] = 0.0;
#pragma unroll 1
  for (int d = 0; d <= D; ++d) {
    const double* SC = mats + (size_t)(3 * d + 0) * msz;
    const double* Mm = mats + (size_t)(3 * d + 1) * msz;
    const double* SK = mats + (size_t)(3 * d + 2) * msz;
    double c0, c1;
    if (d > 0) {  // C(d-1)
      const double* Mp = Mm - 3 * msz;
      double* fgp = S.FG() + S.vix(d - 1, 0, 0);
      const double* nm = d < D ? SC : next_mats;
      mma_task<false, kTr, kTr>(a, trp(Mp, 0), r0, trp(Mp, 1), r1, S.Wb(), nullptr, ns, np, c0, c1);
      cur[0][0] -= c0;
      cur[0][1] -= c1;
      mma_task<false, kTr, kFwd>(a, trp(Mp, 1), r1, nm ? stream_ptr<kFwd>(nm, np, own.b[0], lane) : nullptr, r0,
                                 S.Wb(), nullptr, ns, np, c0, c1);
      cur[1][0] -= c0;
      cur[1][1] -= c1;
      fgp[o[0][0]] = gv[0][0];
      fgp[o[0][1]] = gv[0][1];
      if (has1) {
        fgp[o[1][0]] = gv[1][0];
        fgp[o[1][1]] = gv[1][1];
      }
      // push the finished component: pushed at index D-1 now, shifted D-d more times -> ends at index d-1
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int k = 0; k + 1 < D; ++k) {
          gxr[h][k][0] = gxr[h][k + 1][0];
          gxr[h][k][1] = gxr[h][k + 1][1];
        }
        gxr[h][D - 1][0] = cur[h][0];
        gxr[h][D - 1][1] = cur[h][1];
      }
    }
    if (d < D) {  // A(d)
      const double* xc = S.Xc() + S.vix(d, 0, 0);
      mma_task<false, kFwd, kFwd>(a, fwd(SC, 0), r0, fwd(Mm, 0), r0, xc, nullptr, ns, np, c0, c1);
      cur[0][0] = 2.0 * c0;
      cur[0][1] = 2.0 * c1;
      t1[0] = fma(xc[o[0][0]], c0, t1[0]);
      t1[1] = fma(xc[o[0][1]], c1, t1[1]);
      mma_task<false, kFwd, kFwd>(a, fwd(Mm, 0), r0, fwd(SC, 1), r1, xc, nullptr, ns, np, c0, c1);
      S.Wa()[o[0][0]] = c0;
      S.Wa()[o[0][1]] = c1;
      mma_task<false, kFwd, kFwd>(a, fwd(SC, 1), r1, fwd(Mm, 1), r1, xc, nullptr, ns, np, c0, c1);
      cur[1][0] = 2.0 * c0;
      cur[1][1] = 2.0 * c1;
      if (has1) {
        t1[0] = fma(xc[o[1][0]], c0, t1[0]);
        t1[1] = fma(xc[o[1][1]], c1, t1[1]);
      }
      mma_task<false, kFwd, kFwd>(a, fwd(Mm, 1), r1, fwd(SK, 0), r0, xc, nullptr, ns, np, c0, c1);
      if (has1) {
        S.Wa()[o[1][0]] = c0;
        S.Wa()[o[1][1]] = c1;
      }
    }
    MAGI_PHASE_SYNC();
    if (d < D) {  // B(d)
      const double* fg = S.FG() + S.vix(d, 0, 0);
      mma_task<true, kFwd, kFwd>(a, fwd(SK, 0), r0, fwd(SK, 1), r1, fg, S.Wa(), ns, np, c0, c1);
      gv[0][0] = 2.0 * c0;
      gv[0][1] = 2.0 * c1;
      t2[0] = fma(fg[o[0][0]] - S.Wa()[o[0][0]], c0, t2[0]);
      t2[1] = fma(fg[o[0][1]] - S.Wa()[o[0][1]], c1, t2[1]);
      S.Wb()[o[0][0]] = gv[0][0];
      S.Wb()[o[0][1]] = gv[0][1];
      mma_task<true, kFwd, kTr>(a, fwd(SK, 1), r1, trp(Mm, 0), r0, fg, S.Wa(), ns, np, c0, c1);
      if (has1) {
        gv[1][0] = 2.0 * c0;
        gv[1][1] = 2.0 * c1;
        t2[0] = fma(fg[o[1][0]] - S.Wa()[o[1][0]], c0, t2[0]);
        t2[1] = fma(fg[o[1][1]] - S.Wa()[o[1][1]], c1, t2[1]);
        S.Wb()[o[1][0]] = gv[1][0];
        S.Wb()[o[1][1]] = gv[1][1];
      }
      MAGI_PHASE_SYNC();
    }
  }

  // pointwise epilogue at the own elements: ODE Jacobian terms, likelihood, assemble the gradient
  double red[2][NRED];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int k = 0; k < NRED; ++k) red[q][k] = 0.0;
    red[q][0] = t1[q];
    red[q][1] = t2[q];
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = own.j[h];
    const bool valid = (h == 0 || has1) && j < n;
    const int js = valid ? j : 0;
    double yv[D];
    bool mk[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      mk[d] = valid && mb[(size_t)js * D + d] != 0;
      yv[d] = mk[d] ? yb[(size_t)js * D + d] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = c2 + q;
      double th[P], x[D], gg[D], vx[D], vth[P];
#pragma unroll
      for (int k = 0; k < P; ++k) th[k] = S.th()[k * kCh + ch];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        x[d] = S.Xc()[S.vix(d, ch, js)] + S.mu()[d];
        gg[d] = valid ? S.FG()[S.vix(d, ch, js)] : 0.0;
      }
      M::vjp(x, th, gg, vx, vth);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double e = mk[d] ? x[d] - yv[d] : 0.0;
        red[q][2 + d] = fma(e, e, red[q][2 + d]);
        gxr[h][d][q] = -0.5 * ((gxr[h][d][q] + vx[d]) * inv_beta + 2.0 * e / S.sig2()[d * kCh + ch]);
      }
#pragma unroll
      for (int k = 0; k < P; ++k) red[q][2 + D + k] += valid ? vth[k] : 0.0;
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      const double v = fold_g(red[q][k]);
      if (lane < 4) S.wpart()[((size_t)warp * kCh + c2 + q) * NRED + k] = v;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int e = tid, nthr = 32 * nw; e < kCh * NRED; e += nthr) {  // (chain, k) totals over the warps in a fixed order
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

// Per-item loads: dataset constants and the chain states (reference layout X[n][D] per chain) into shared
// memory at the own elements, centred; chains >= nr and grid indices >= n are zero.  Ends with
// __syncthreads().
template <class M, int NP>
__device__ void fast_load_item(const FastScratch<M, NP>& S, const magi_problem_t& pb, int b, const double* X,
                               const double* sig_pre, const double* th_pre, size_t chain0, int nr) {
  constexpr int D = M::D, P = M::P;
  const int n = S.n, tid = threadIdx.x, nthr = 32 * S.nw();
  const Own own = fast_own(S);
  double mu[D];
#pragma unroll
  for (int d = 0; d < D; ++d) mu[d] = pb.mu[(size_t)b * D + d];
  if (tid < D) {
    S.mu()[tid] = pb.mu[(size_t)b * D + tid];
    S.Nd()[tid] = pb.N_ds[(size_t)b * D + tid];
    S.LB()[tid] = pb.LB[(size_t)b * D + tid];
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (own.b[h] >= 0) {
      const int j = own.j[h];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int ch = own.c2 + q;
        const bool ok = ch < nr && j < n;
        const double* xp = X + ((chain0 + (ok ? ch : 0)) * n + (ok ? j : 0)) * D;
#pragma unroll
        for (int d = 0; d < D; ++d) S.Xc()[S.vix(d, ch, j)] = ok ? xp[d] - mu[d] : 0.0;
      }
    }
  }
  for (int e = tid; e < kCh * D; e += nthr) {
    const int r = e / D, d = e - r * D;
    S.s()[d * kCh + r] = r < nr ? sig_pre[(chain0 + r) * D + d] : 0.0;
  }
  for (int e = tid; e < kCh * P; e += nthr) {
    const int r = e / P, k = e - r * P;
    S.tau()[k * kCh + r] = r < nr ? th_pre[(chain0 + r) * P + k] : 0.0;
  }
  __syncthreads();
}
