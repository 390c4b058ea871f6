// Fast path of the posterior core for the headline shapes: np = 8*ceil(n/8) <= 168 (n <= 168 grid
// points, e.g. the reference's SEIR setting n = 161) and D <= 4 components.
//
// Same math and the same DMMA contractions as posterior_core.cuh; what differs is how the matrices reach the
// tensor cores and where everything else lives:
//   * one warp owns one 8-row block of every matrix; lane = 4g + c owns grid index j = 8*warp + g of
//     chains 2c and 2c+1 ("own elements") in ALL phases, pointwise ones included;
//   * the matrix tiles are TMA-STAGED: every warp has a private ring of 2 ... 3 chunks (slots) x kTsC tiles (512 B each) in
//     shared memory, filled by cp.async.bulk copies (tma_stage.cuh) that complete on per-chunk mbarriers.  The
//     sequence of tiles a warp needs -- pass after pass, evaluation after evaluation, item after item -- does not
//     depend on any computed value, so lane 0 of the warp re-arms a chunk with the tiles one ring length ahead the
//     moment the chunk has been consumed (TileStream): the bytes in flight are bounded by the ring, not by
//     registers or L1 lines, and they keep flowing across __syncthreads, the pointwise epilogue and item changes;
//   * banded matrices (tf.linalg.band_part, magi_v2.py:271-274): a warp streams only the tiles of its block row /
//     column that can hold non-zeros (11 ... 21 of 21 at n = 161, band 80: 75 % of the bytes).  The ring makes the
//     unequal trip counts harmless -- the operands come from shared memory by address, not from a register file
//     indexed at compile time -- and the phases are paced by the arrival of bytes, not by the longest warp;
//   * the gradient dL/dX lives in registers (2*D doubles per lane) from the first contraction to the
//     final store / momentum kick;
//   * momentum and the saved start point of an HMC transition live in a per-CTA global scratch slot
//     in own-element order (coalesced 16-byte accesses, L2-resident because the grid is persistent);
//   * shared memory: the rings (fast_ring_plan: ~95 KB), Xc, FG [D][8][ns], Wa, Wb [8][ns], y/mask,
//     per-chain scalars (~124 KB at n = 161).
// The grid is persistent (one CTA per SM looping over (dataset, chain-group) items).
#pragma once
#include "posterior_core.cuh"
#include "tma_stage.cuh"

constexpr int kFastMaxNp = 168;
constexpr int kFastMaxD = 4;
constexpr int kFastMaxBlk0 = kFastMaxNp / 8;
#ifndef MAGI_TS_PROD
#define MAGI_TS_PROD 3   // producer warps (copy issue only); consumer warp w is served by lane w / kTsProd of producer w % kTsProd
#endif
constexpr int kTsProd = MAGI_TS_PROD;
constexpr int kFastMaxThreads = 32 * (kFastMaxBlk0 + kTsProd);
#ifndef MAGI_TS_CPP
#define MAGI_TS_CPP 3   // chunks (= mbarrier phases = bulk copies of a forward pass) a warp's block row is cut into
#endif
constexpr int kTsCpp = MAGI_TS_CPP;
#ifndef MAGI_TS_SLOTS
#define MAGI_TS_SLOTS 2
#endif
constexpr int kTsSlots = MAGI_TS_SLOTS;   // chunks per warp ring
constexpr int kTsMaxC = 7;          // largest chunk, tiles (its fragments are held in registers: 4 per tile)
constexpr int kFastSmemBytes = 232448;   // 227 KB: the most dynamic shared memory a CTA can have on sm_100

// How the ring space is shared out.  Warp w cuts the `len` tiles of its block row into chunks of
// cw = min(ceil(len / kTsCpp), cap) tiles and owns a ring of kTsSlots such chunks starting at tile `off`: with banded
// matrices the middle rows are up to twice as long as the edge rows, so every warp keeps the same FRACTION of a
// pass in flight and all warps reach the barrier that ends a phase together.  `cap` is the largest chunk size for
// which all rings fit into `budget_tiles`.  Same arithmetic on host and device.
struct RingPlan {
  int off, cw, total;   // tiles
};
__host__ __device__ inline RingPlan fast_ring_plan(int np, int band, int warp, int budget_tiles) {
  const int nblk = np >> 3, kb = band < 0 ? nblk : (band + 7) >> 3;
  RingPlan p{0, 1, 0};
  for (int cap = kTsMaxC; cap >= 1; --cap) {
    p.off = p.total = 0;
    for (int w = 0; w < nblk; ++w) {
      const int lo = w - kb > 0 ? w - kb : 0, hi = w + kb + 1 < nblk ? w + kb + 1 : nblk;
      int cw = (hi - lo + kTsCpp - 1) / kTsCpp;
      if (cw > cap) cw = cap;
      if (w < warp) p.off += kTsSlots * cw;
      if (w == warp) p.cw = cw;
      p.total += kTsSlots * cw;
    }
    if (p.total <= budget_tiles) break;
  }
  return p;
}

// NP > 0: the padded grid size is a compile-time constant (all shared-memory offsets and trip counts fold
// into immediates -- the instantiation the headline n = 161 runs); NP = 0: run-time np <= kFastMaxNp.
template <class M, int NP>
struct FastScratch {
  static constexpr int NRED = 2 + M::D + M::P;            // t1, t2, SSE_d, sum_j vth_k
  static constexpr int kCtl = 10 + 2 * M::P + 2 * M::D;   // control rows of the HMC kernel (sampler_fast.cuh)
  static constexpr int kSmall = kCh * (5 * M::P + 6 * M::D + 1 + kCtl) + 3 * M::D;
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
  __device__ __forceinline__ double* ctl() const { return mu() + 3 * M::D; }           // [kCtl][8]
  __device__ __forceinline__ double* isig2() const { return ctl() + kCtl * kCh; }      // [D][8] 1 / sigma^2

  // doubles of everything but the rings
  __host__ __device__ static size_t fixed_elems(int np_) {
    const int ns_ = magi_chain_stride(np_);
    const size_t e = (size_t)2 * M::D * kCh * ns_ + (size_t)2 * kCh * ns_ + kSmall;
    return (e + 1) & ~(size_t)1;
  }
  __host__ __device__ static RingPlan ring_plan(int np_, int band, int warp) {
    const long left = (long)kFastSmemBytes - (long)(fixed_elems(np_) + ring_bar_elems(np_)) * 8;
    return fast_ring_plan(np_, band, warp, (int)(left / 512));
  }
  __host__ __device__ static size_t ring_bar_elems(int np_) { return (size_t)(np_ >> 3) * kTsSlots * 2; }   // full | empty
  // doubles at the start of dynamic shared memory taken by the rings (tiles of 64 doubles) and their mbarriers
  __host__ __device__ static size_t ring_elems(int np_, int band) {
    return (size_t)ring_plan(np_, band, 0).total * 64 + ring_bar_elems(np_);
  }
  __host__ __device__ static size_t elems(int np_, int band) { return fixed_elems(np_) + ring_elems(np_, band); }
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


// ---- optional timeline instrumentation (tools/trace_fast.py; -DMAGI_TRACE builds only) ----------------------------
#ifdef MAGI_TRACE
constexpr int kTraceCap = 8192;
__device__ unsigned long long g_magi_trace[kFastMaxBlk0 * kTraceCap];   // [warp][event]: (clock64 << 8) | tag
__device__ int g_magi_trace_n[kFastMaxBlk0];
#define MAGI_TR(tag)                                                                                          \
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ts.tri < kTraceCap) {                                     \
    g_magi_trace[(threadIdx.x >> 5) * kTraceCap + ts.tri] = ((unsigned long long)clock64() << 8) | (tag);     \
    g_magi_trace_n[threadIdx.x >> 5] = ++ts.tri;                                                              \
  }
#else
#define MAGI_TR(tag)
#endif

// ---- items of the persistent grid ----------------------------------------------------------------
template <class M>
__device__ __forceinline__ bool fast_item(const magi_problem_t& pb, int item, int& b, int& nr, size_t& chain0) {
  const int groups = (pb.R + kCh - 1) / kCh;
  if (item >= pb.B * groups) return false;
  b = item / groups;
  const int r0 = (item - b * groups) * kCh;
  nr = min(kCh, pb.R - r0);
  chain0 = (size_t)b * pb.R + r0;
  return true;
}

template <class M>
__device__ __forceinline__ const double* fast_mats(const magi_problem_t& pb, int b, int np) {
  return static_cast<const double*>(pb.packed) + (size_t)b * M::D * 3 * np * np;
}

// ---- TMA-staged tile stream of one warp ------------------------------------------------------------
// One evaluation is 4 D matrix passes, always in this order (fast_eval consumes them in the same order):
//   pass p = 4 d + kind:   kind 0: m_d forward | 1: sym(K_d^-1) forward | 2: m_d transposed | 3: sym(C_d^-1) forward
// and in every pass warp w needs the tiles t in [lo, hi) of block row w (forward; contiguous in the tiled layout:
// ONE bulk copy per chunk) or of block column w (transposed: one 512-byte copy per tile, from L2 -- the forward read
// of the same matrix was two passes earlier).  Chunks never straddle passes.  The producer side (p*) runs one ring
// length ahead of the consumer side (k); both walk the same sequence, so no tags are needed.
// consumer side of a warp's stream
template <class M>
struct TileStream {
  static constexpr int kPasses = 4 * M::D;
  uint32_t ring, bars;   // shared-space addresses: this warp's ring (kTsSlots chunks of cw tiles), its mbarriers
                         // (kTsSlots "full" ones, then kTsSlots "empty" ones)
  int lo, len, cw;       // tile steps [lo, lo + len) of every pass, cut into chunks of cw tiles
  uint32_t k;            // chunks consumed (slot = k % kTsSlots, mbarrier parity = (k / kTsSlots) & 1)
#ifdef MAGI_TRACE
  int tri;               // events recorded so far
#endif
};

// ring geometry of consumer warp `warp` (shared by both sides)
template <class M, int NP>
__device__ __forceinline__ void ts_geometry(const FastScratch<M, NP>& S, int band, int warp, uint32_t& ring, uint32_t& bars,
                                            int& lo, int& len, int& cw) {
  const int nblk = S.nblk();
  const RingPlan rp = FastScratch<M, NP>::ring_plan(S.np(), band, warp);
  double* ring0 = S.base - FastScratch<M, NP>::ring_elems(S.np(), band);
  ring = smem_u32(ring0 + (size_t)rp.off * 64);
  bars = smem_u32(ring0 + (size_t)rp.total * 64 + warp * kTsSlots * 2);
  cw = rp.cw;
  const StepRange r = band_range(warp, nblk, band_blocks(band, nblk));
  lo = r.lo;
  len = r.hi - r.lo;
}

// Consumer warps: set up the stream state and initialise this warp's mbarriers.  Must be followed by a
// __syncthreads() of the WHOLE CTA before ts_producer starts (fast_kernel_prologue does both).
template <class M, int NP>
__device__ __forceinline__ void ts_init(const FastScratch<M, NP>& S, TileStream<M>& ts, const magi_problem_t& pb) {
  ts_geometry(S, pb.band, threadIdx.x >> 5, ts.ring, ts.bars, ts.lo, ts.len, ts.cw);
  ts.k = 0;
#ifdef MAGI_TRACE
  ts.tri = 0;
#endif
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 2 * kTsSlots; ++q) mbar_init(ts.bars + 8 * q, 1);
    mbar_init_fence();
  }
}

// matrices of item `item` of the persistent grid, or nullptr beyond the last item
template <class M>
__device__ __forceinline__ const double* ts_item_mats(const magi_problem_t& pb, int item, int np) {
  int b, nr;
  size_t c0;
  return fast_item<M>(pb, item, b, nr, c0) ? fast_mats<M>(pb, b, np) : nullptr;
}

// Producer warps (threadIdx.x >= 32 * nblk): lane l of producer warp q issues the bulk copies of consumer warp
// w = q + kTsProd * l, chunk after chunk, for every evaluation of every item of this CTA: it waits until the
// consumer has handed the slot back ("empty" mbarrier), arms the slot's "full" mbarrier with the chunk's byte count
// and issues the copies.  The lanes poll without blocking, so that one slow consumer does not hold up the others.
// `evals_per_item`: how many times fast_eval runs for every item (1: log-posterior; n_steps + 1: leapfrog;
// 1 + n_iter * n_leapfrog: HMC).  Returns when the streams are exhausted.
template <class M, int NP>
__device__ void ts_producer(const FastScratch<M, NP>& S, const magi_problem_t& pb, int evals_per_item) {
  const int nblk = S.nblk(), np = S.np(), lane = threadIdx.x & 31;
  const int w = ((int)(threadIdx.x >> 5) - nblk) + kTsProd * lane;   // the consumer warp this lane serves
  uint32_t ring, bars;
  int lo, len, cw;
  bool active = w < nblk;
  ts_geometry(S, pb.band, active ? w : 0, ring, bars, lo, len, cw);
  const ptrdiff_t msz = (ptrdiff_t)np * np;
  const ptrdiff_t fwd = ((ptrdiff_t)w * nblk + lo) * 64, tr = ((ptrdiff_t)lo * nblk + w) * 64;
  int item = blockIdx.x, pev = evals_per_item, ppass = 0, pleft = len;
  uint32_t kp = 0;   // chunks issued
  // L2 eviction priorities of the copies (measured: 1.94 -> 1.86 ms): the tiles of m are read again two passes later
  // and should still be in L2 then; everything else passes through once per evaluation and should not push them out
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  const double* mats = ts_item_mats<M>(pb, item, np);
  const double* psrc = mats ? mats + msz + fwd : nullptr;   // pass 0: m of component 0, forward
  active = active && psrc != nullptr;
  while (__any_sync(MAGI_FULL_MASK, active)) {
    const uint32_t slot = kp % kTsSlots;
    // the first kTsSlots chunks go into fresh slots; afterwards wait for the consumer's (kp / kTsSlots)-th release
    const bool go = active && (kp < kTsSlots || mbar_test(bars + 8 * (kTsSlots + slot), ((kp / kTsSlots) - 1) & 1));
    if (!__any_sync(MAGI_FULL_MASK, go)) {   // nothing to issue for any lane: leave the issue slots of this scheduler to
      __nanosleep(100);                      // its consumer warps for a moment (measured 20 / 100 / 300 / 1000 ns: +0.3 /
      continue;                              // +0.4 / +0.3 / 0 %)
    }
    if (go) {
      const int nt = min(cw, pleft);
      const int kind = ppass & 3;
      const uint32_t bar = bars + 8 * slot, dst = ring + slot * (uint32_t)(cw * 512);
      mbar_arrive_expect_tx(bar, (uint32_t)nt * 512);
      if (kind != 2) {
        bulk_g2s_hint(dst, psrc, (uint32_t)nt * 512, bar, kind == 0 ? pol_keep : pol_stream);
        psrc += (ptrdiff_t)nt * 64;
      } else {
#pragma unroll 1
        for (int i = 0; i < nt; ++i) {
          bulk_g2s_hint(dst + i * 512, psrc, 512, bar, pol_stream);
          psrc += (ptrdiff_t)nblk * 64;
        }
      }
      ++kp;
      pleft -= nt;
      if (pleft == 0) {   // next pass / evaluation / item
        pleft = len;
        // from the end of the finished pass to the first tile of the next one:
        //   kind 0 (m fwd) -> 1 (S_K fwd): + one matrix;  1 -> 2 (m tr): back to m, row / column swapped;
        //   2 -> 3 (S_C fwd): back to S_C;  3 -> 0 of d+1 (m fwd): + 4 matrices
        const double* p0 = psrc - (ptrdiff_t)len * (kind == 2 ? nblk * 64 : 64);
        if (++ppass < 4 * M::D) {
          psrc = kind == 0 ? p0 + msz : kind == 1 ? p0 - msz - fwd + tr : kind == 2 ? p0 - msz - tr + fwd : p0 + 4 * msz;
        } else {
          ppass = 0;
          if (--pev == 0) {
            item += gridDim.x;
            pev = evals_per_item;
            mats = ts_item_mats<M>(pb, item, np);
          }
          psrc = mats ? mats + msz + fwd : nullptr;
          active = psrc != nullptr;
        }
      }
    }
  }
}

// One matrix pass of this warp (the next one of the stream): y = A x over the warp's tile steps, CT = kFwd for a
// block row, kTr for a block column (the transposed fragment is read directly: the swizzled tile layout of common.cuh
// makes both fragment shapes conflict-free); SUB: the vector is x - xsub.
// Result: c0, c1 = y[chain 2c], y[chain 2c+1] at block element g (lane = 4g + c).
// (Measured before the swizzle, profiles/r02_headline.md: transposing in registers with six shuffles per tile, lazily or
// for a whole chunk; contracting a transposed pass with the tile as the B operand.)
template <class M, int NP, bool SUB, int CT>
__device__ __forceinline__ void staged_pass(const FastScratch<M, NP>& S, TileStream<M>& ts, const double* x,
                                            const double* xsub, double& c0, double& c1) {
  const int lane = threadIdx.x & 31, ns = S.ns();
  const size_t bo = (size_t)(lane >> 2) * ns + 2 * (lane & 3) + 8 * ts.lo;
  const double* bp = x + bo;
  const double* bs = SUB ? xsub + bo : nullptr;
  // byte offsets of this lane's fragment inside a (swizzled, common.cuh) tile: forward = the pair (g, c); transposed =
  // the elements (2c, g) and (2c+1, g), i.e. exactly the fragment of the transposed tile -- no shuffles
  const int fg = lane >> 2, fc = lane & 3;
  const uint32_t ofw = 16 * magi_tile_slot(fg, fc);
  const uint32_t otr0 = 8 * magi_tile_pos(2 * fc, fg), otr1 = 8 * magi_tile_pos(2 * fc + 1, fg);
  double acc[2][2];  // two independent accumulator pairs (even / odd k-group of a step)
#pragma unroll
  for (int q = 0; q < 2; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 1
  for (int left = ts.len; left > 0; left -= ts.cw) {
    const int nt = min(ts.cw, left);
    const uint32_t slot = ts.k % kTsSlots;
    MAGI_TR(2)
#ifndef MAGI_DIAG_NOLOAD
    mbar_wait(ts.bars + 8 * slot, (ts.k / kTsSlots) & 1);
#endif
    MAGI_TR(3)
    const uint32_t ap = ts.ring + slot * (uint32_t)(ts.cw * 512);
    // The chunk's matrix fragments go to registers first and the slot is handed back to the producer AT ONCE (the
    // refill is in flight while this chunk is contracted).  All kTsMaxC loads are unconditional: beyond the chunk's
    // last tile they read other, finite ring contents that are never used.
    double2 a[kTsMaxC];
#pragma unroll
    for (int i = 0; i < kTsMaxC; ++i) {
      if (CT == kTr) a[i] = make_double2(lds_f64(ap + i * 512 + otr0), lds_f64(ap + i * 512 + otr1));
      else a[i] = lds_f64x2(ap + i * 512 + ofw);
    }
    __syncwarp();   // every lane has read the slot
    ++ts.k;
    if (lane == 0) mbar_arrive(ts.bars + 8 * (kTsSlots + slot));
    MAGI_TR(4)
    // the nt tiles as straight-line code entered through a warp-uniform switch (tile nt-1 first, falling through
    // to tile 0): no predicates, no loop counters, every offset an immediate
#define MAGI_TILE(i)                                                            \
  {                                                                             \
    const double2 av = a[i];                                                    \
    double2 b = *reinterpret_cast<const double2*>(bp + 8 * (i));                \
    if (SUB) {                                                                  \
      const double2 b2 = *reinterpret_cast<const double2*>(bs + 8 * (i));       \
      b.x -= b2.x;                                                              \
      b.y -= b2.y;                                                              \
    }                                                                           \
    dmma(acc[0][0], acc[0][1], av.x, b.x);                                      \
    dmma(acc[1][0], acc[1][1], av.y, b.y);                                      \
  }
#ifndef MAGI_DIAG_NOCOMPUTE
    static_assert(kTsMaxC == 7, "the switch below is written for chunks of up to 7 tiles");
    switch (nt) {
      case 7: MAGI_TILE(6)
      case 6: MAGI_TILE(5)
      case 5: MAGI_TILE(4)
      case 4: MAGI_TILE(3)
      case 3: MAGI_TILE(2)
      case 2: MAGI_TILE(1)
      default: MAGI_TILE(0)
    }
#endif
#undef MAGI_TILE
    bp += 8 * nt;
    if (SUB) bs += 8 * nt;
    MAGI_TR(5)
  }
  c0 = acc[0][0] + acc[1][0];
  c1 = acc[0][1] + acc[1][1];
}

// Evaluate base log-posterior L and gradient at the state in shared memory (S.Xc(), S.tau(), S.s()).
// Consumes the next 4 D passes of the warp's tile stream.  On exit S.L(), S.gs(), S.gtau() are set and
// gxr[d][q] = dL/dX[j, d] of chain 2c+q at the lane's own grid index (garbage for j >= n).  All consumer threads must call.
template <class M, int NP>
__device__ void fast_eval(const FastScratch<M, NP>& S, TileStream<M>& ts, const magi_problem_t& pb, int b,
                          double inv_beta, double (&gxr)[M::D][2]) {
  constexpr int D = M::D, P = M::P, NRED = FastScratch<M, NP>::NRED;
  const int n = S.n, np = S.np(), ns = S.ns();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = S.nblk();
  const int g = lane >> 2, c2 = 2 * (lane & 3);
  const int j = warp * 8 + g;
  const bool valid = j < n;
  const size_t o0 = (size_t)c2 * ns + j, o1 = o0 + ns;  // own elements inside one [8][ns] array
  MAGI_TR(10)

  // per-chain transforms of the small state parts (magi_v2.py:318-319): one thread per (parameter, chain)
#pragma unroll 1
  for (int e = tid, nthr = 32 * nw; e < (P + D) * kCh; e += nthr) {
    const int q = e >> 3, c = e & 7;
    if (q < P) {
      const double t = S.tau()[q * kCh + c];
      S.th()[q * kCh + c] = magi_softplus(t);
      S.sgt()[q * kCh + c] = magi_sigmoid(t);
    } else {
      const int d = q - P;
      const double z = S.s()[d * kCh + c];
      const double s2 = magi_softplus(z) + S.LB()[d];
      S.sig2()[d * kCh + c] = s2;
      S.isig2()[d * kCh + c] = 1.0 / s2;
      S.sgs()[d * kCh + c] = magi_sigmoid(z);
    }
  }
  named_sync(32 * S.nblk());
  MAGI_TR(11)

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
  // (no barrier: the first pass below reads Xc, which is complete, and writes Wa; f is first read by the second pass)
  MAGI_TR(12)

  double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
  // Pass order per component d (= the order of the tile stream):
  //   [w = m_d xc -> Wa] | [g = 2 S_K (f - w) -> Wb] | [v = m_d^T g ; u = S_C xc ; w of d+1]      ("|" = __syncthreads)
  auto pass_m = [&](int d) {
    double c0, c1;
    staged_pass<M, NP, false, kFwd>(S, ts, S.Xc() + S.vix(d, 0, 0), nullptr, c0, c1);
    S.Wa()[o0] = c0;
    S.Wa()[o1] = c1;
  };
  pass_m(0);
  named_sync(32 * S.nblk());
  MAGI_TR(13)
#pragma unroll 1
  for (int d = 0; d < D; ++d) {
    const double* xc = S.Xc() + S.vix(d, 0, 0);
    double* fg = S.FG() + S.vix(d, 0, 0);
    double g0, g1;
    {  // g = 2 S_K (f - w) -> Wb ; t2 += (f - w) . S_K (f - w)
      double c0, c1;
      staged_pass<M, NP, true, kFwd>(S, ts, fg, S.Wa(), c0, c1);
      g0 = 2.0 * c0;
      g1 = 2.0 * c1;
      S.Wb()[o0] = g0;
      S.Wb()[o1] = g1;
      t2[0] = fma(fg[o0] - S.Wa()[o0], c0, t2[0]);
      t2[1] = fma(fg[o1] - S.Wa()[o1], c1, t2[1]);
    }
    named_sync(32 * S.nblk());
    MAGI_TR(14)
    {  // v = m^T g (second read of m: L2) ; u = S_C xc ; gxr = 2u - v ; FG_d <- g ; w of d+1
      double v0, v1, u0, u1;
      if (d == D - 1 && valid) {   // the epilogue's observation loads: have their lines in L2 by then
        l2_prefetch_keep(pb.y + ((size_t)b * n + j) * D);
        l2_prefetch_keep(pb.mask + ((size_t)b * n + j) * D);
      }
      staged_pass<M, NP, false, kTr>(S, ts, S.Wb(), nullptr, v0, v1);
      staged_pass<M, NP, false, kFwd>(S, ts, xc, nullptr, u0, u1);
      t1[0] = fma(xc[o0], u0, t1[0]);
      t1[1] = fma(xc[o1], u1, t1[1]);
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {   // (static register indices)
        if (dd == d) {
          gxr[dd][0] = 2.0 * u0 - v0;
          gxr[dd][1] = 2.0 * u1 - v1;
        }
      }
      fg[o0] = g0;
      fg[o1] = g1;
      if (d + 1 < D) pass_m(d + 1);
    }
    named_sync(32 * S.nblk());
    MAGI_TR(15)
  }

  // pointwise epilogue at the own elements: ODE Jacobian terms, likelihood, assemble the gradient.
  // The observations stay in global memory (there is no shared memory left for them): all loads of this lane are
  // issued together here, and their lines were prefetched into L2 during the last component (below).
  double yv[D];
  bool obs[D];
  {
    const size_t ob = ((size_t)b * n + (valid ? j : 0)) * D;
    if (D == 4 && (reinterpret_cast<uintptr_t>(pb.y) & 15) == 0 && (reinterpret_cast<uintptr_t>(pb.mask) & 3) == 0) {
      // 32 bytes of y and 4 bytes of mask per grid point
      const double2 y01 = ldg_f64x2(pb.y + ob), y23 = ldg_f64x2(pb.y + ob + 2);
      const unsigned mk = __ldg(reinterpret_cast<const unsigned*>(pb.mask + ob));
      yv[0] = y01.x; yv[1] = y01.y; yv[2 % D] = y23.x; yv[3 % D] = y23.y;
#pragma unroll
      for (int d = 0; d < D; ++d) obs[d] = valid && ((mk >> (8 * d)) & 0xffu) != 0;
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        obs[d] = valid && pb.mask[ob + d] != 0;
        yv[d] = pb.y[ob + d];
      }
    }
  }
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
      const double e = obs[d] ? x[d] - yv[d] : 0.0;
      red[2 + d] = e * e;
      gxr[d][q] = -0.5 * ((gxr[d][q] + vx[d]) * inv_beta + 2.0 * e * S.isig2()[d * kCh + ch]);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) red[2 + D + k] = valid ? vth[k] : 0.0;
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      const double v = fold_g(red[k]);
      if (lane < 4) S.wpart()[((size_t)warp * kCh + ch) * NRED + k] = v;
    }
  }
  named_sync(32 * S.nblk());
  MAGI_TR(16)
  // Totals over the warps (fixed order) and the per-chain results: 16 lanes per chain, lane q < D: component q
  // (noise term, Jacobian of the softplus, dL/ds), q - D < P: parameter (dL/dtau), q = 15: the two quadratic forms;
  // the 16 contributions to L are added by shuffles.
  static_assert(D + P <= 15, "one lane per component / parameter plus one for the quadratic forms");
#pragma unroll 1
  for (int e = tid, nthr = 32 * nw; e < kCh * 16; e += nthr) {
    const int c = e >> 4, q = e & 15;
    const int col = q < D + P ? 2 + q : 0;
    double v = 0.0;
    if (q < D + P || q == 15) {
      const double* wp = S.wpart() + (size_t)c * NRED + col;
#pragma unroll 1
      for (int w = 0; w < nw; ++w) {
        v += wp[(size_t)w * kCh * NRED];
        if (q == 15) v += wp[(size_t)w * kCh * NRED + 1];
      }
    }
    double lc = 0.0;   // contribution to L
    if (q < D) {
      const double s2 = S.sig2()[q * kCh + c], sg = S.sgs()[q * kCh + c], z = S.s()[q * kCh + c], Nd = S.Nd()[q];
      const double r = v / s2;
      lc = -0.5 * (Nd * log(2.0 * M_PI * s2) + r) + (z - magi_softplus(z));
      S.gs()[q * kCh + c] = -0.5 * (Nd / s2 - r / s2) * sg + (1.0 - sg);
    } else if (q < D + P) {
      const int k = q - D;
      const double t = S.tau()[k * kCh + c], sg = S.sgt()[k * kCh + c];
      lc = t - magi_softplus(t);
      S.gtau()[k * kCh + c] = -0.5 * inv_beta * v * sg + (1.0 - sg);
    } else if (q == 15) {
      lc = -0.5 * inv_beta * v;
    }
    lc += magi_shfl_xor(lc, 8);
    lc += magi_shfl_xor(lc, 4);
    lc += magi_shfl_xor(lc, 2);
    lc += magi_shfl_xor(lc, 1);
    if (q == 0) S.L()[c] = lc;
  }
  // (a barrier of the four warps of this stage alone -- everybody else going straight on to the gradient stores or the
  // kick of X -- was measured: no gain, profiles/r02_headline.md)
  named_sync(32 * S.nblk());
  MAGI_TR(17)
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
      double xv[D];
      if (D == 4 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {   // 32 contiguous bytes: two 128-bit loads
        const double2 x01 = ldg_f64x2(xp), x23 = ldg_f64x2(xp + 2);
        xv[0] = x01.x; xv[1] = x01.y; xv[2 % D] = x23.x; xv[3 % D] = x23.y;
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) xv[d] = xp[d];
      }
#pragma unroll
      for (int d = 0; d < D; ++d) S.Xc()[S.vix(d, ch, j)] = ok ? xv[d] - mu[d] : 0.0;
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
  named_sync(32 * S.nblk());
}
