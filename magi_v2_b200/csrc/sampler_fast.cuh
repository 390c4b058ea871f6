// Fast-path kernels (np <= 168, D <= 4; see posterior_fast.cuh): persistent grid, one CTA per SM
// looping over (dataset, chain-group) items; gradient in registers, momentum / saved start point in a
// per-CTA global scratch slot.  Included by sampler.cu (inside its anonymous namespace).
#pragma once
#include "posterior_fast.cuh"
#include "rng.cuh"

struct HmcOut {
  double *th_samps, *sig_samps, *X_samps, *X_sum, *X_sumsq, *accept_prob, *lp_trace;
};

template <class M, int NP>
__device__ __forceinline__ void fast_setup(FastScratch<M, NP>& S, int n, int band) {
  extern __shared__ __align__(128) double smem[];
  S.base = smem + FastScratch<M, NP>::ring_elems(magi_pad8(n), band);   // the tile rings come first (posterior_fast.cuh)
  S.n = n;
  S.np_rt = magi_pad8(n);
}

// Common start of the warp-specialised fast kernels: the consumer warps (threadIdx.x < 32 * nblk) set up their tile
// streams; after a barrier of the whole CTA the producer warps run the copy loop and return true (the caller returns).
template <class M, int NP>
__device__ __forceinline__ bool fast_kernel_prologue(const FastScratch<M, NP>& S, TileStream<M>& ts, const magi_problem_t& pb,
                                                     int evals_per_item) {
  const bool producer = (int)(threadIdx.x >> 5) >= S.nblk();
  if (!producer) ts_init(S, ts, pb);
  __syncthreads();
  if (producer) {
#ifndef MAGI_DIAG_NOLOAD
    ts_producer(S, pb, evals_per_item);
#endif
  }
  return producer;
}

// ---- (3b) log-posterior + gradient ---------------------------------------------------------------
template <class M, int NP>
__global__ void __launch_bounds__(kFastMaxThreads, 1)
logpost_grad_fast_kernel(magi_problem_t pb, const double* __restrict__ X, const double* __restrict__ sig_pre,
                         const double* __restrict__ th_pre, const double* __restrict__ beta_temp,
                         double* __restrict__ lp, double* __restrict__ gX, double* __restrict__ gsig,
                         double* __restrict__ gth) {
  constexpr int D = M::D, P = M::P;
  FastScratch<M, NP> S;
  fast_setup<M, NP>(S, pb.n, pb.band);
  const int n = S.n, np = S.np(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c2 = 2 * (lane & 3), j = warp * 8 + g;
  int b, nr;
  size_t chain0;
  TileStream<M> ts;
  if (fast_kernel_prologue(S, ts, pb, 1)) return;
  for (int item = blockIdx.x; fast_item<M>(pb, item, b, nr, chain0); item += gridDim.x) {
    fast_load_item(S, pb, b, X, sig_pre, th_pre, chain0, nr);
    MAGI_TR(21)
    {  // the next item's chain states (own elements): into L2 while this item is evaluated
      int bn, nrn;
      size_t c0n;
      if (fast_item<M>(pb, item + gridDim.x, bn, nrn, c0n) && j < n) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (c2 + q < nrn) l2_prefetch_keep(X + ((c0n + c2 + q) * n + j) * D);
      }
      if (tid == 0) l2_prefetch_keep(beta_temp + chain0);   // read after the evaluation
    }
    const double* btv = beta_temp + chain0;
    double gxr[D][2];
    fast_eval(S, ts, pb, b, 1.0 / pb.beta[b], gxr);
    // scale by the temperature and store in the reference layout X[n][D]
    if (j < n) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (c2 + q < nr) {
          const double bt = btv[c2 + q];
          double* o = gX + ((chain0 + c2 + q) * n + j) * D;
          if (D == 4 && (reinterpret_cast<uintptr_t>(gX) & 15) == 0) {   // two 128-bit stores
            *reinterpret_cast<double2*>(o) = make_double2(bt * gxr[0][q], bt * gxr[1][q]);
            *reinterpret_cast<double2*>(o + 2) = make_double2(bt * gxr[2 % D][q], bt * gxr[3 % D][q]);
          } else {
#pragma unroll
            for (int d = 0; d < D; ++d) o[d] = bt * gxr[d][q];
          }
        }
      }
    }
    MAGI_TR(20)
    if (tid < nr) {
      const double bt = btv[tid];
      lp[chain0 + tid] = bt * S.L()[tid];
#pragma unroll
      for (int d = 0; d < D; ++d) gsig[(chain0 + tid) * D + d] = bt * S.gs()[d * kCh + tid];
#pragma unroll
      for (int k = 0; k < P; ++k) gth[(chain0 + tid) * P + k] = bt * S.gtau()[k * kCh + tid];
    }
    // no barrier here: fast_eval ended with one, the stores above read registers and S.L / S.gs / S.gtau only, and
    // the next item's loads touch none of those before the barriers of its own evaluation
  }
}

// ---- leapfrog machinery ----------------------------------------------------------------------------
// p += ck * eps * bt * grad  and, if drift, z += eps * p  -- on all three state parts.  The momentum of
// X is in the CTA's global scratch slot (own-element order), the gradient of X in registers.
template <class M, int NP>
__device__ __forceinline__ void fast_kick_drift(const FastScratch<M, NP>& S, double* PX, const double (&gxr)[M::D][2],
                                                const double* epsv, const double* btv, double ck, bool drift) {
  constexpr int D = M::D, P = M::P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c2 = 2 * (lane & 3), j = warp * 8 + g;
  const int nblk = S.np() >> 3;
  if (j < S.n) {
    const double e0 = epsv[c2], e1 = epsv[c2 + 1];
    const double h0 = ck * e0 * btv[c2], h1 = ck * e1 * btv[c2 + 1];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double2* pp = reinterpret_cast<double2*>(PX + own_ix(d, nblk));
      double2 p = *pp;
      p.x = fma(h0, gxr[d][0], p.x);
      p.y = fma(h1, gxr[d][1], p.y);
      *pp = p;
      if (drift) {
        const size_t i0 = S.vix(d, c2, j), i1 = i0 + S.ns();
        S.Xc()[i0] = fma(e0, p.x, S.Xc()[i0]);
        S.Xc()[i1] = fma(e1, p.y, S.Xc()[i1]);
      }
    }
  }
  // small state parts: one thread per (parameter, chain) -- the SAME thread that transforms that parameter at the
  // start of fast_eval, which therefore needs no barrier in between
#pragma unroll 1
  for (int e = tid, nthr = 32 * S.nblk(); e < (P + D) * kCh; e += nthr) {
    const int q = e >> 3, c = e & 7;
    const double hh = ck * epsv[c] * btv[c], ee = epsv[c];
    if (q < P) {
      const double p = fma(hh, S.gtau()[q * kCh + c], S.ptau()[q * kCh + c]);
      S.ptau()[q * kCh + c] = p;
      if (drift) S.tau()[q * kCh + c] = fma(ee, p, S.tau()[q * kCh + c]);
    } else {
      const int d = q - P;
      const double p = fma(hh, S.gs()[d * kCh + c], S.ps()[d * kCh + c]);
      S.ps()[d * kCh + c] = p;
      if (drift) S.s()[d * kCh + c] = fma(ee, p, S.s()[d * kCh + c]);
    }
  }
}

// TFP SimpleLeapfrogIntegrator: per step  p += eps/2 g;  z += eps p;  g = grad(z);  p += eps/2 g  (the two
// half kicks of consecutive steps are applied as one).  Needs the gradient at the current z in gxr / S.gs() /
// S.gtau() on entry; leaves the gradient at the end point.
template <class M, int NP>
__device__ void fast_leapfrog_steps(const FastScratch<M, NP>& S, TileStream<M>& ts, const magi_problem_t& pb, int b,
                                    double* PX, double (&gxr)[M::D][2], double inv_beta, const double* epsv,
                                    const double* btv, int n_steps) {
  if (n_steps <= 0) return;
  fast_kick_drift(S, PX, gxr, epsv, btv, 0.5, true);
  for (int st = 0; st < n_steps; ++st) {
    const bool last = st + 1 == n_steps;
    fast_eval(S, ts, pb, b, inv_beta, gxr);  // starts with a __syncthreads-protected phase
    fast_kick_drift(S, PX, gxr, epsv, btv, last ? 0.5 : 1.0, !last);
  }
  named_sync(32 * S.nblk());
}

// out[r] = 1/2 |p_r|^2 over all three parts (fixed summation order).  Uses S.wpart() / S.tot().
template <class M, int NP>
__device__ void fast_kinetic(const FastScratch<M, NP>& S, const double* PX, double* out) {
  constexpr int D = M::D, P = M::P, NRED = FastScratch<M, NP>::NRED;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = S.nblk();
  const int g = lane >> 2, c2 = 2 * (lane & 3), j = warp * 8 + g;
  const int nblk = S.nblk();
  double k0 = 0.0, k1 = 0.0;
  if (j < S.n) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double2 p = *reinterpret_cast<const double2*>(PX + own_ix(d, nblk));
      k0 = fma(p.x, p.x, k0);
      k1 = fma(p.y, p.y, k1);
    }
  }
  k0 = fold_g(k0);
  k1 = fold_g(k1);
  if (lane < 4) {
    S.wpart()[((size_t)warp * kCh + c2) * NRED] = k0;
    S.wpart()[((size_t)warp * kCh + c2 + 1) * NRED] = k1;
  }
  named_sync(32 * S.nblk());
  if (tid < kCh) {
    double t = 0.0;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) t += S.wpart()[((size_t)w * kCh + tid) * NRED];
#pragma unroll
    for (int d = 0; d < D; ++d) t = fma(S.ps()[d * kCh + tid], S.ps()[d * kCh + tid], t);
#pragma unroll
    for (int k = 0; k < P; ++k) t = fma(S.ptau()[k * kCh + tid], S.ptau()[k * kCh + tid], t);
    out[tid] = 0.5 * t;
  }
  named_sync(32 * S.nblk());
}

// ---- (3c) leapfrog with caller-supplied momenta ---------------------------------------------------
template <class M, int NP>
__global__ void __launch_bounds__(kFastMaxThreads, 1)
leapfrog_fast_kernel(magi_problem_t pb, double* X, double* sig_pre, double* th_pre, double* pX, double* psig,
                     double* pth, const double* __restrict__ eps, const double* __restrict__ beta_temp, int n_steps,
                     double* lp_out, double* ws) {
  constexpr int D = M::D, P = M::P;
  FastScratch<M, NP> S;
  fast_setup<M, NP>(S, pb.n, pb.band);
  const int n = S.n, np = S.np(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c2 = 2 * (lane & 3), j = warp * 8 + g, nblk = np >> 3;
  double* PX = ws + (size_t)blockIdx.x * fast_slot_elems<M>(np);
  double* epsv = S.ctl();
  double* btv = S.ctl() + kCh;
  int b, nr;
  size_t chain0;
  TileStream<M> ts;
  if (fast_kernel_prologue(S, ts, pb, max(n_steps, 0) + 1)) return;
  for (int item = blockIdx.x; fast_item<M>(pb, item, b, nr, chain0); item += gridDim.x) {
    fast_load_item(S, pb, b, X, sig_pre, th_pre, chain0, nr);
    // momenta into own-element order
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double2 p = make_double2(0.0, 0.0);
      if (j < n) {
        if (c2 < nr) p.x = pX[((chain0 + c2) * n + j) * D + d];
        if (c2 + 1 < nr) p.y = pX[((chain0 + c2 + 1) * n + j) * D + d];
      }
      *reinterpret_cast<double2*>(PX + own_ix(d, nblk)) = p;
    }
    if (tid < kCh) {
      const bool ok = tid < nr;
      epsv[tid] = ok ? eps[chain0 + tid] : 0.0;
      btv[tid] = ok ? beta_temp[chain0 + tid] : 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) S.ps()[d * kCh + tid] = ok ? psig[(chain0 + tid) * D + d] : 0.0;
#pragma unroll
      for (int k = 0; k < P; ++k) S.ptau()[k * kCh + tid] = ok ? pth[(chain0 + tid) * P + k] : 0.0;
    }
    named_sync(32 * S.nblk());
    const double inv_beta = 1.0 / pb.beta[b];
    double gxr[D][2];
    fast_eval(S, ts, pb, b, inv_beta, gxr);
    fast_leapfrog_steps(S, ts, pb, b, PX, gxr, inv_beta, epsv, btv, n_steps);

    if (j < n) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (c2 + q < nr) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const size_t o = ((chain0 + c2 + q) * n + j) * D + d;
            X[o] = S.Xc()[S.vix(d, c2 + q, j)] + S.mu()[d];
            const double2 p = *reinterpret_cast<const double2*>(PX + own_ix(d, nblk));
            pX[o] = q ? p.y : p.x;
          }
        }
      }
    }
    if (tid < nr) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        sig_pre[(chain0 + tid) * D + d] = S.s()[d * kCh + tid];
        psig[(chain0 + tid) * D + d] = S.ps()[d * kCh + tid];
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        th_pre[(chain0 + tid) * P + k] = S.tau()[k * kCh + tid];
        pth[(chain0 + tid) * P + k] = S.ptau()[k * kCh + tid];
      }
      if (lp_out) lp_out[chain0 + tid] = btv[tid] * S.L()[tid];
    }
    named_sync(32 * S.nblk());
  }
}

// ---- (3d) HMC sampler: all iterations of a group of 8 chains inside one CTA -------------------------
template <class M, int NP>
__global__ void __launch_bounds__(kFastMaxThreads, 1)
hmc_fast_kernel(magi_problem_t pb, magi_hmc_config_t cfg, double* X, double* sig_pre, double* th_pre, double* eps,
                double* da_state, HmcOut out, double* ws) {
  constexpr int D = M::D, P = M::P;
  FastScratch<M, NP> S;
  fast_setup<M, NP>(S, pb.n, pb.band);
  const int n = S.n, np = S.np(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c2 = 2 * (lane & 3), j = warp * 8 + g, nblk = np >> 3;
  const size_t slot = fast_slot_elems<M>(np);
  double* PX = ws + (size_t)blockIdx.x * 3 * slot;   // momentum
  double* X0 = PX + slot;                            // start point of the transition (centred)
  double* G0 = X0 + slot;                            // its gradient
  const size_t nchains = (size_t)pb.B * pb.R;
  // control block in shared memory
  double* epsv = S.ctl();               // [8]
  double* btv = S.ctl() + 1 * kCh;      // [8]
  double* ke = S.ctl() + 2 * kCh;       // [8] scratch for kinetic energies
  double* h0 = S.ctl() + 3 * kCh;       // [8]
  double* accf = S.ctl() + 4 * kCh;     // [8] 1.0 = accepted
  double* L0 = S.ctl() + 5 * kCh;       // [8]
  double* da = S.ctl() + 6 * kCh;       // [4][8]
  double* tau0 = S.ctl() + 10 * kCh;    // [P][8]
  double* gtau0 = tau0 + P * kCh;     // [P][8]
  double* s0 = gtau0 + P * kCh;       // [D][8]
  double* gs0 = s0 + D * kCh;         // [D][8]
  const int nstate = n * D + D + P;

  int b, nr;
  size_t chain0;
  TileStream<M> ts;
  if (fast_kernel_prologue(S, ts, pb, 1 + max(cfg.n_iter, 0) * max(cfg.n_leapfrog, 0))) return;
  for (int item = blockIdx.x; fast_item<M>(pb, item, b, nr, chain0); item += gridDim.x) {
    fast_load_item(S, pb, b, X, sig_pre, th_pre, chain0, nr);
    if (tid < kCh) {
      const bool ok = tid < nr;
      epsv[tid] = ok ? eps[chain0 + tid] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) da[q * kCh + tid] = ok ? da_state[(chain0 + tid) * 4 + q] : 0.0;
    }
    named_sync(32 * S.nblk());
    const double inv_beta = 1.0 / pb.beta[b];
    double gxr[D][2];
    fast_eval(S, ts, pb, b, inv_beta, gxr);

    for (int it = 0; it < cfg.n_iter; ++it) {
      const int git = cfg.iter0 + it;
      const double bt = cfg.fixed_beta_temp > 0.0 ? cfg.fixed_beta_temp
                                                  : fmax(1.0 / log((double)git + 2.0), cfg.min_temp);
      // save the start point and its gradient; draw momenta ~ N(0, I): element e of the packed state
      // (X row-major [n][D], then s, then tau) is normal number e of the chain's Philox stream
      if (j < n) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const size_t o = own_ix(d, nblk);
          *reinterpret_cast<double2*>(X0 + o) = make_double2(S.Xc()[S.vix(d, c2, j)], S.Xc()[S.vix(d, c2 + 1, j)]);
          *reinterpret_cast<double2*>(G0 + o) = make_double2(gxr[d][0], gxr[d][1]);
        }
        double pq[D][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t cid = cfg.chain_id0 + (uint32_t)(chain0 + c2 + q);
          const int e0 = j * D;
#pragma unroll
          for (int d = 0; d < D; ++d) pq[d][q] = 0.0;
          if (c2 + q < nr) {
            // pairs covering elements e0 .. e0 + D - 1
            for (int pr = e0 >> 1; pr <= (e0 + D - 1) >> 1; ++pr) {
              double z0, z1;
              magi_normal_pair(cfg.seed, (uint32_t)pr, cid, (uint32_t)git, z0, z1);
#pragma unroll
              for (int d = 0; d < D; ++d) {
                if (2 * pr == e0 + d) pq[d][q] = z0;
                if (2 * pr + 1 == e0 + d) pq[d][q] = z1;
              }
            }
          }
        }
#pragma unroll
        for (int d = 0; d < D; ++d)
          *reinterpret_cast<double2*>(PX + own_ix(d, nblk)) = make_double2(pq[d][0], pq[d][1]);
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) *reinterpret_cast<double2*>(PX + own_ix(d, nblk)) = make_double2(0.0, 0.0);
      }
      if (tid < kCh) {
        btv[tid] = bt;
        L0[tid] = S.L()[tid];
#pragma unroll
        for (int d = 0; d < D; ++d) { s0[d * kCh + tid] = S.s()[d * kCh + tid]; gs0[d * kCh + tid] = S.gs()[d * kCh + tid]; }
#pragma unroll
        for (int k = 0; k < P; ++k) { tau0[k * kCh + tid] = S.tau()[k * kCh + tid]; gtau0[k * kCh + tid] = S.gtau()[k * kCh + tid]; }
        // momenta of the small parts: elements n*D .. n*D + D + P - 1
        if (tid < nr) {
          const uint32_t cid = cfg.chain_id0 + (uint32_t)(chain0 + tid);
          for (int e = n * D; e < nstate; ++e) {
            double z0, z1;
            magi_normal_pair(cfg.seed, (uint32_t)(e >> 1), cid, (uint32_t)git, z0, z1);
            const double z = (e & 1) ? z1 : z0;
            if (e < n * D + D) S.ps()[(e - n * D) * kCh + tid] = z;
            else S.ptau()[(e - n * D - D) * kCh + tid] = z;
          }
        } else {
#pragma unroll
          for (int d = 0; d < D; ++d) S.ps()[d * kCh + tid] = 0.0;
#pragma unroll
          for (int k = 0; k < P; ++k) S.ptau()[k * kCh + tid] = 0.0;
        }
      }
      named_sync(32 * S.nblk());
      fast_kinetic(S, PX, ke);
      if (tid < kCh) h0[tid] = -bt * S.L()[tid] + ke[tid];
      named_sync(32 * S.nblk());

      fast_leapfrog_steps(S, ts, pb, b, PX, gxr, inv_beta, epsv, btv, cfg.n_leapfrog);

      fast_kinetic(S, PX, ke);
      if (tid < kCh) {
        const double h1 = -bt * S.L()[tid] + ke[tid];
        const double dH = h1 - h0[tid];
        double ap = 0.0;
        if (isfinite(dH)) ap = fmin(1.0, exp(fmin(0.0, -dH)));
        bool accepted = false;
        if (tid < nr) {
          const double u = magi_uniform(cfg.seed, cfg.chain_id0 + (uint32_t)(chain0 + tid), (uint32_t)git);
          accepted = u < ap;
        }
        accf[tid] = accepted ? 1.0 : 0.0;
        if (!accepted) {
          S.L()[tid] = L0[tid];
#pragma unroll
          for (int d = 0; d < D; ++d) { S.s()[d * kCh + tid] = s0[d * kCh + tid]; S.gs()[d * kCh + tid] = gs0[d * kCh + tid]; }
#pragma unroll
          for (int k = 0; k < P; ++k) { S.tau()[k * kCh + tid] = tau0[k * kCh + tid]; S.gtau()[k * kCh + tid] = gtau0[k * kCh + tid]; }
        }
        // dual averaging (tfp DualAveragingStepSizeAdaptation restated; oracle: dual_averaging_update)
        const double step = da[3 * kCh + tid];
        if (step < (double)cfg.num_adapt) {
          const double err = da[0 * kCh + tid] + (cfg.target_accept - ap);
          const double t = step + 1.0;
          const double log_x = da[2 * kCh + tid] - sqrt(t) * err / (0.05 * (t + 10.0));
          const double eta = pow(t, -0.75);
          const double lavg = eta * log_x + (1.0 - eta) * da[1 * kCh + tid];
          da[0 * kCh + tid] = err;
          da[1 * kCh + tid] = lavg;
          epsv[tid] = (step + 1.0 == (double)cfg.num_adapt) ? exp(lavg) : exp(log_x);
        }
        da[3 * kCh + tid] = step + 1.0;
        if (tid < nr) {
          const size_t o = (size_t)it * nchains + chain0 + tid;
          if (out.accept_prob) out.accept_prob[o] = ap;
          if (out.lp_trace) out.lp_trace[o] = bt * S.L()[tid];
          if (out.th_samps) {
#pragma unroll
            for (int k = 0; k < P; ++k) out.th_samps[o * P + k] = magi_softplus(S.tau()[k * kCh + tid]);
          }
          if (out.sig_samps) {
#pragma unroll
            for (int d = 0; d < D; ++d) out.sig_samps[o * D + d] = magi_softplus(S.s()[d * kCh + tid]) + S.LB()[d];
          }
        }
      }
      named_sync(32 * S.nblk());
      // rejected chains go back to the start point; then emit the trajectory sample
      if (j < n) {
        const bool acc0 = accf[c2] != 0.0, acc1 = accf[c2 + 1] != 0.0;
        const bool accum = git >= cfg.accum_from && (out.X_sum || out.X_sumsq);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const size_t o = own_ix(d, nblk);
          if (!acc0 || !acc1) {
            const double2 x0 = *reinterpret_cast<const double2*>(X0 + o);
            const double2 g0 = *reinterpret_cast<const double2*>(G0 + o);
            if (!acc0) { S.Xc()[S.vix(d, c2, j)] = x0.x; gxr[d][0] = g0.x; }
            if (!acc1) { S.Xc()[S.vix(d, c2 + 1, j)] = x0.y; gxr[d][1] = g0.y; }
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (c2 + q < nr) {
              const double xv = S.Xc()[S.vix(d, c2 + q, j)] + S.mu()[d];
              const size_t go = ((chain0 + c2 + q) * n + j) * D + d;
              if (out.X_samps) out.X_samps[(size_t)it * nchains * n * D + go] = xv;
              if (accum) {
                if (out.X_sum) out.X_sum[go] += xv;
                if (out.X_sumsq) out.X_sumsq[go] = fma(xv, xv, out.X_sumsq[go]);
              }
            }
          }
        }
      }
      named_sync(32 * S.nblk());
    }

    // write back the chain state
    if (j < n) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (c2 + q < nr) {
#pragma unroll
          for (int d = 0; d < D; ++d)
            X[((chain0 + c2 + q) * n + j) * D + d] = S.Xc()[S.vix(d, c2 + q, j)] + S.mu()[d];
        }
      }
    }
    if (tid < nr) {
#pragma unroll
      for (int d = 0; d < D; ++d) sig_pre[(chain0 + tid) * D + d] = S.s()[d * kCh + tid];
#pragma unroll
      for (int k = 0; k < P; ++k) th_pre[(chain0 + tid) * P + k] = S.tau()[k * kCh + tid];
      eps[chain0 + tid] = epsv[tid];
#pragma unroll
      for (int q = 0; q < 4; ++q) da_state[(chain0 + tid) * 4 + q] = da[q * kCh + tid];
    }
    named_sync(32 * S.nblk());
  }
}
