// Kernels and C-ABI entry points for (3a)-(3d) of include/magi_b200.h:
// pack_matrices, logpost_grad, leapfrog, hmc_run.
#include "posterior_core.cuh"
#include "rng.cuh"

namespace {

#include "sampler_fast.cuh"

constexpr size_t kMaxSmem = 227 * 1024;

// ------------------------------------------------------------------------------------------------
// pack: [B,D,n,n] x3 (reference layout) -> [B][D][3][np/8][np/8][8][8] tiles with symmetrised quadratic forms
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const double* __restrict__ Cinv, const double* __restrict__ m,
                            const double* __restrict__ Kinv, int n, int np, double* __restrict__ out) {
  const size_t bd = blockIdx.z;
  const int i = blockIdx.y;
  const double* c = Cinv + bd * (size_t)n * n;
  const double* mm = m + bd * (size_t)n * n;
  const double* k = Kinv + bd * (size_t)n * n;
  double* o = out + bd * 3 * (size_t)np * np;
  const int nblk = np >> 3;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < np; j += gridDim.x * blockDim.x) {
    double sc = 0.0, mv = 0.0, sk = 0.0;
    if (i < n && j < n) {
      sc = 0.5 * (c[(size_t)i * n + j] + c[(size_t)j * n + i]);
      mv = mm[(size_t)i * n + j];
      sk = 0.5 * (k[(size_t)i * n + j] + k[(size_t)j * n + i]);
    }
    const size_t t = ((size_t)(i >> 3) * nblk + (j >> 3)) * 64 + magi_tile_pos(i & 7, j & 7);  // tiled, swizzled (common.cuh)
    o[t] = sc;
    o[(size_t)np * np + t] = mv;
    o[2 * (size_t)np * np + t] = sk;
  }
}

// ------------------------------------------------------------------------------------------------
// scratch placement: big vector arrays in shared memory when they fit, else in the workspace
// ------------------------------------------------------------------------------------------------
// one warp per 8-row block of the matrices, between 8 and 21 warps
inline int threads_for(int n) {
  const int nblk = magi_pad8(n) / 8;
  const int nw = nblk < kMinWarps ? kMinWarps : (nblk > kMaxWarps ? kMaxWarps : nblk);
  return 32 * nw;
}

template <class M>
struct Placement {
  size_t smem_bytes;      // dynamic shared memory per CTA
  size_t ws_big_elems;    // per-CTA doubles of global scratch for the big arrays (0 if in smem)
  size_t ws_save_elems;   // per-CTA doubles of global save area (HMC: z0 and grad0)
  bool big_in_smem;
};

template <class M>
Placement<M> placement(int n, bool with_momentum, bool with_save) {
  const int np = magi_pad8(n);
  const size_t big = Scratch<M>::big_elems(np, with_momentum) * sizeof(double);
  const size_t small = Scratch<M>::small_elems(np) * sizeof(double);
  Placement<M> p;
  p.big_in_smem = big + small <= kMaxSmem;
  p.smem_bytes = p.big_in_smem ? big + small : small;
  p.ws_big_elems = p.big_in_smem ? 0 : Scratch<M>::big_elems(np, with_momentum);
  p.ws_save_elems = with_save ? (size_t)2 * M::D * kCh * magi_chain_stride(np) : 0;
  return p;
}

// SMEM = true: the vector arrays are carved out of dynamic shared memory (compile-time known, so the
// compiler emits LDS/STS rather than generic loads); false: they live in the caller's workspace.
template <class M, bool SMEM>
__device__ __forceinline__ void setup_scratch(Scratch<M>& S, double* ws, size_t ws_big, size_t ws_save, int n,
                                              bool with_momentum, double*& save) {
  extern __shared__ __align__(128) double smem[];
  const int np = magi_pad8(n);
  const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  double* wsc = ws ? ws + cta * (ws_big + ws_save) : nullptr;
  if (SMEM) {
    S.carve(smem + Scratch<M>::small_elems(np), smem, n, np, with_momentum);
    S.assume_shared();
  } else {
    S.carve(wsc, smem, n, np, with_momentum);
  }
  save = ws_save ? wsc + ws_big : nullptr;
}

// ------------------------------------------------------------------------------------------------
// (3b) log-posterior + gradient
// ------------------------------------------------------------------------------------------------
template <class M, bool SMEM>
__global__ void __launch_bounds__(kMaxThreads, 1)
logpost_grad_kernel(magi_problem_t pb, const double* __restrict__ X, const double* __restrict__ sig_pre,
                    const double* __restrict__ th_pre, const double* __restrict__ beta_temp,
                    double* __restrict__ lp, double* __restrict__ gX, double* __restrict__ gsig,
                    double* __restrict__ gth, double* ws, size_t ws_big) {
  constexpr int D = M::D, P = M::P;
  Scratch<M> S;
  double* save;
  setup_scratch<M, SMEM>(S, ws, ws_big, 0, pb.n, false, save);
  const int b = blockIdx.y, r0 = blockIdx.x * kCh;
  const int nr = min(kCh, pb.R - r0);
  const size_t chain0 = (size_t)b * pb.R + r0;
  const int n = S.n, np = S.np, tid = threadIdx.x;
  load_dataset(S, pb, b);
  __syncthreads();
  load_state(S, X, sig_pre, th_pre, chain0, nr);
  __syncthreads();
  const double* mats = static_cast<const double*>(pb.packed) + (size_t)b * D * 3 * np * np;
  eval_logpost_grad(S, mats, 1.0 / pb.beta[b], pb.band);

  // scale by the temperature and store in the reference layout
  const int per = n * D;
  for (int e = tid; e < nr * per; e += blockDim.x) {
    const int r = e / per, rem = e - r * per;
    const int j = rem / D, d = rem - j * D;
    gX[(chain0 + r) * per + rem] = beta_temp[chain0 + r] * S.GX[S.vix(d, r, j)];
  }
  if (tid < nr) {
    const double bt = beta_temp[chain0 + tid];
    lp[chain0 + tid] = bt * S.L[tid];
#pragma unroll
    for (int d = 0; d < D; ++d) gsig[(chain0 + tid) * D + d] = bt * S.gs[d * kCh + tid];
#pragma unroll
    for (int k = 0; k < P; ++k) gth[(chain0 + tid) * P + k] = bt * S.gtau[k * kCh + tid];
  }
}

// ------------------------------------------------------------------------------------------------
// leapfrog machinery shared by (3c) and (3d)
// ------------------------------------------------------------------------------------------------
// p += c * eps * bt * grad   on all three state parts.  epsv/btv: [8] in shared memory.
template <class M>
__device__ __forceinline__ void kick(const Scratch<M>& S, const double* epsv, const double* btv, double c) {
  constexpr int D = M::D, P = M::P;
  const int tid = threadIdx.x;
  const ChainMap cm = chain_map();
  const int r = cm.r;
  const double h = c * epsv[r] * btv[r];
  for (int j = cm.j0; j < S.n; j += cm.jstride) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const size_t a = S.vix(d, r, j);
      S.PX[a] = fma(h, S.GX[a], S.PX[a]);
    }
  }
  if (tid < kCh) {
    const double hh = c * epsv[tid] * btv[tid];
#pragma unroll
    for (int d = 0; d < D; ++d) S.ps[d * kCh + tid] = fma(hh, S.gs[d * kCh + tid], S.ps[d * kCh + tid]);
#pragma unroll
    for (int k = 0; k < P; ++k) S.ptau[k * kCh + tid] = fma(hh, S.gtau[k * kCh + tid], S.ptau[k * kCh + tid]);
  }
}

// z += eps * p
template <class M>
__device__ __forceinline__ void drift(const Scratch<M>& S, const double* epsv) {
  constexpr int D = M::D, P = M::P;
  const int tid = threadIdx.x;
  const ChainMap cm = chain_map();
  const int r = cm.r;
  const double h = epsv[r];
  for (int j = cm.j0; j < S.n; j += cm.jstride) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const size_t a = S.vix(d, r, j);
      S.Xc[a] = fma(h, S.PX[a], S.Xc[a]);
    }
  }
  if (tid < kCh) {
    const double hh = epsv[tid];
#pragma unroll
    for (int d = 0; d < D; ++d) S.s[d * kCh + tid] = fma(hh, S.ps[d * kCh + tid], S.s[d * kCh + tid]);
#pragma unroll
    for (int k = 0; k < P; ++k) S.tau[k * kCh + tid] = fma(hh, S.ptau[k * kCh + tid], S.tau[k * kCh + tid]);
  }
}

// TFP SimpleLeapfrogIntegrator: per step  p += eps/2 g;  z += eps p;  g = grad(z);  p += eps/2 g.
// Requires the gradient at the current z in scratch on entry; leaves the gradient at the end point.
template <class M>
__device__ void leapfrog_steps(const Scratch<M>& S, const double* mats, double inv_beta, int band,
                               const double* epsv, const double* btv, int n_steps) {
  for (int st = 0; st < n_steps; ++st) {
    kick(S, epsv, btv, 0.5);
    __syncthreads();
    drift(S, epsv);
    __syncthreads();
    eval_logpost_grad(S, mats, inv_beta, band);
    kick(S, epsv, btv, 0.5);
    __syncthreads();
  }
}

// out[r] = 1/2 |p_r|^2 over all three parts (deterministic two-level reduction).
template <class M>
__device__ void kinetic(const Scratch<M>& S, double* out) {
  constexpr int D = M::D, P = M::P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const ChainMap cm = chain_map();
  const int r = cm.r;
  double acc = 0.0;
  for (int j = cm.j0; j < S.n; j += cm.jstride) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double p = S.PX[S.vix(d, r, j)];
      acc = fma(p, p, acc);
    }
  }
  acc = magi_warp_sum(acc);
  if (lane == 0) S.wpart[warp] = acc;
  __syncthreads();
  if (tid < kCh) {
    double t = 0.0;
    for (int w = tid; w < nw; w += 8) t += S.wpart[w];
#pragma unroll
    for (int d = 0; d < D; ++d) t = fma(S.ps[d * kCh + tid], S.ps[d * kCh + tid], t);
#pragma unroll
    for (int k = 0; k < P; ++k) t = fma(S.ptau[k * kCh + tid], S.ptau[k * kCh + tid], t);
    out[tid] = 0.5 * t;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// (3c) leapfrog with caller-supplied momenta
// ------------------------------------------------------------------------------------------------
template <class M, bool SMEM>
__global__ void __launch_bounds__(kMaxThreads, 1)
leapfrog_kernel(magi_problem_t pb, double* X, double* sig_pre, double* th_pre, double* pX, double* psig,
                double* pth, const double* __restrict__ eps, const double* __restrict__ beta_temp, int n_steps,
                double* lp_out, double* ws, size_t ws_big) {
  constexpr int D = M::D, P = M::P;
  Scratch<M> S;
  double* save;
  setup_scratch<M, SMEM>(S, ws, ws_big, 0, pb.n, true, save);
  const int b = blockIdx.y, r0 = blockIdx.x * kCh;
  const int nr = min(kCh, pb.R - r0);
  const size_t chain0 = (size_t)b * pb.R + r0;
  const int n = S.n, np = S.np, tid = threadIdx.x;
  double* epsv = S.ctl;
  double* btv = S.ctl + kCh;
  load_dataset(S, pb, b);
  __syncthreads();
  load_state(S, X, sig_pre, th_pre, chain0, nr);
  const int per = n * D;
  for (size_t e = tid; e < S.vsize(); e += blockDim.x) S.PX[e] = 0.0;
  __syncthreads();
  for (int e = tid; e < nr * per; e += blockDim.x) {
    const int r = e / per, rem = e - r * per;
    const int j = rem / D, d = rem - j * D;
    S.PX[S.vix(d, r, j)] = pX[(chain0 + r) * per + rem];
  }
  if (tid < kCh) {
    const bool ok = tid < nr;
    epsv[tid] = ok ? eps[chain0 + tid] : 0.0;
    btv[tid] = ok ? beta_temp[chain0 + tid] : 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) S.ps[d * kCh + tid] = ok ? psig[(chain0 + tid) * D + d] : 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) S.ptau[k * kCh + tid] = ok ? pth[(chain0 + tid) * P + k] : 0.0;
  }
  __syncthreads();
  const double* mats = static_cast<const double*>(pb.packed) + (size_t)b * D * 3 * np * np;
  const double inv_beta = 1.0 / pb.beta[b];
  eval_logpost_grad(S, mats, inv_beta, pb.band);
  leapfrog_steps(S, mats, inv_beta, pb.band, epsv, btv, n_steps);

  for (int e = tid; e < nr * per; e += blockDim.x) {
    const int r = e / per, rem = e - r * per;
    const int j = rem / D, d = rem - j * D;
    X[(chain0 + r) * per + rem] = S.Xc[S.vix(d, r, j)] + S.mu[d];
    pX[(chain0 + r) * per + rem] = S.PX[S.vix(d, r, j)];
  }
  if (tid < nr) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      sig_pre[(chain0 + tid) * D + d] = S.s[d * kCh + tid];
      psig[(chain0 + tid) * D + d] = S.ps[d * kCh + tid];
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
      th_pre[(chain0 + tid) * P + k] = S.tau[k * kCh + tid];
      pth[(chain0 + tid) * P + k] = S.ptau[k * kCh + tid];
    }
    if (lp_out) lp_out[chain0 + tid] = btv[tid] * S.L[tid];
  }
}

// ------------------------------------------------------------------------------------------------
// (3d) HMC sampler: all iterations of a group of 8 chains in one launch
// ------------------------------------------------------------------------------------------------
template <class M, bool SMEM>
__global__ void __launch_bounds__(kMaxThreads, 1)
hmc_kernel(magi_problem_t pb, magi_hmc_config_t cfg, double* X, double* sig_pre, double* th_pre, double* eps,
           double* da_state, HmcOut out, double* ws, size_t ws_big, size_t ws_save) {
  constexpr int D = M::D, P = M::P;
  Scratch<M> S;
  double* save;
  setup_scratch<M, SMEM>(S, ws, ws_big, ws_save, pb.n, true, save);
  const int b = blockIdx.y, r0 = blockIdx.x * kCh;
  const int nr = min(kCh, pb.R - r0);
  const size_t chain0 = (size_t)b * pb.R + r0;
  const size_t nchains = (size_t)pb.B * pb.R;
  const int n = S.n, np = S.np, tid = threadIdx.x;
  const ChainMap cm = chain_map();
  const int r = cm.r;
  const size_t v = S.vsize();
  // control block in shared memory
  double* epsv = S.ctl;               // [8]
  double* btv = S.ctl + 1 * kCh;      // [8]
  double* ke = S.ctl + 2 * kCh;       // [8] scratch for kinetic energies
  double* h0 = S.ctl + 3 * kCh;       // [8]
  double* accf = S.ctl + 4 * kCh;     // [8] 1.0 = accepted
  double* L0 = S.ctl + 5 * kCh;       // [8]
  double* da = S.ctl + 6 * kCh;       // [4][8]
  double* tau0 = S.ctl + 16 * kCh;    // [P][8]
  double* gtau0 = tau0 + P * kCh;     // [P][8]
  double* s0 = gtau0 + P * kCh;       // [D][8]
  double* gs0 = s0 + D * kCh;         // [D][8]
  double* X0 = save;                  // [D][8][np] global
  double* G0 = save + v;

  load_dataset(S, pb, b);
  __syncthreads();
  load_state(S, X, sig_pre, th_pre, chain0, nr);
  if (tid < kCh) {
    const bool ok = tid < nr;
    epsv[tid] = ok ? eps[chain0 + tid] : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) da[q * kCh + tid] = ok ? da_state[(chain0 + tid) * 4 + q] : 0.0;
  }
  __syncthreads();
  const double* mats = static_cast<const double*>(pb.packed) + (size_t)b * D * 3 * np * np;
  const double inv_beta = 1.0 / pb.beta[b];
  eval_logpost_grad(S, mats, inv_beta, pb.band);

  const int nstate = n * D + D + P;
  const int npairs = (nstate + 1) >> 1;
  for (int it = 0; it < cfg.n_iter; ++it) {
    const int git = cfg.iter0 + it;
    const double bt = cfg.fixed_beta_temp > 0.0 ? cfg.fixed_beta_temp
                                                : fmax(1.0 / log((double)git + 2.0), cfg.min_temp);
    // save the start point and its gradient
    for (int j = cm.j0; j < n; j += cm.jstride) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const size_t a = S.vix(d, r, j);
        X0[a] = S.Xc[a];
        G0[a] = S.GX[a];
      }
    }
    if (tid < kCh) {
      btv[tid] = bt;
      L0[tid] = S.L[tid];
#pragma unroll
      for (int d = 0; d < D; ++d) { s0[d * kCh + tid] = S.s[d * kCh + tid]; gs0[d * kCh + tid] = S.gs[d * kCh + tid]; }
#pragma unroll
      for (int k = 0; k < P; ++k) { tau0[k * kCh + tid] = S.tau[k * kCh + tid]; gtau0[k * kCh + tid] = S.gtau[k * kCh + tid]; }
    }
    // momenta ~ N(0, I): element e of the packed state (X row-major [n][D], then s, then tau)
    if (r < nr) {
      const uint32_t cid = cfg.chain_id0 + (uint32_t)(chain0 + r);
      for (int q = cm.j0; q < npairs; q += cm.jstride) {
        double z[2];
        magi_normal_pair(cfg.seed, (uint32_t)q, cid, (uint32_t)git, z[0], z[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = 2 * q + h;
          if (e < n * D) {
            const int j = e / D, d = e - j * D;
            S.PX[S.vix(d, r, j)] = z[h];
          } else if (e < n * D + D) {
            S.ps[(e - n * D) * kCh + r] = z[h];
          } else if (e < nstate) {
            S.ptau[(e - n * D - D) * kCh + r] = z[h];
          }
        }
      }
    }
    __syncthreads();
    kinetic(S, ke);
    if (tid < kCh) h0[tid] = -bt * S.L[tid] + ke[tid];
    __syncthreads();

    leapfrog_steps(S, mats, inv_beta, pb.band, epsv, btv, cfg.n_leapfrog);

    kinetic(S, ke);
    if (tid < kCh) {
      const double h1 = -bt * S.L[tid] + ke[tid];
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
        S.L[tid] = L0[tid];
#pragma unroll
        for (int d = 0; d < D; ++d) { S.s[d * kCh + tid] = s0[d * kCh + tid]; S.gs[d * kCh + tid] = gs0[d * kCh + tid]; }
#pragma unroll
        for (int k = 0; k < P; ++k) { S.tau[k * kCh + tid] = tau0[k * kCh + tid]; S.gtau[k * kCh + tid] = gtau0[k * kCh + tid]; }
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
        if (out.lp_trace) out.lp_trace[o] = bt * S.L[tid];
        if (out.th_samps) {
#pragma unroll
          for (int k = 0; k < P; ++k) out.th_samps[o * P + k] = magi_softplus(S.tau[k * kCh + tid]);
        }
        if (out.sig_samps) {
#pragma unroll
          for (int d = 0; d < D; ++d) out.sig_samps[o * D + d] = magi_softplus(S.s[d * kCh + tid]) + S.LB[d];
        }
      }
    }
    __syncthreads();
    // rejected chains go back to the start point; then emit the trajectory sample
    const bool acc = accf[r] != 0.0;
    const bool accum = git >= cfg.accum_from && (out.X_sum || out.X_sumsq);
    for (int j = cm.j0; j < n; j += cm.jstride) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const size_t a = S.vix(d, r, j);
        if (!acc) {
          S.Xc[a] = X0[a];
          S.GX[a] = G0[a];
        }
        if (r < nr) {
          const double xv = S.Xc[a] + S.mu[d];
          const size_t g = ((chain0 + r) * n + j) * D + d;
          if (out.X_samps) out.X_samps[(size_t)it * nchains * n * D + g] = xv;
          if (accum) {
            if (out.X_sum) out.X_sum[g] += xv;
            if (out.X_sumsq) out.X_sumsq[g] = fma(xv, xv, out.X_sumsq[g]);
          }
        }
      }
    }
    __syncthreads();
  }

  // write back the chain state
  const int per = n * D;
  for (int e = tid; e < nr * per; e += blockDim.x) {
    const int rr = e / per, rem = e - rr * per;
    const int j = rem / D, d = rem - j * D;
    X[(chain0 + rr) * per + rem] = S.Xc[S.vix(d, rr, j)] + S.mu[d];
  }
  if (tid < nr) {
#pragma unroll
    for (int d = 0; d < D; ++d) sig_pre[(chain0 + tid) * D + d] = S.s[d * kCh + tid];
#pragma unroll
    for (int k = 0; k < P; ++k) th_pre[(chain0 + tid) * P + k] = S.tau[k * kCh + tid];
    eps[chain0 + tid] = epsv[tid];
#pragma unroll
    for (int q = 0; q < 4; ++q) da_state[(chain0 + tid) * 4 + q] = da[q * kCh + tid];
  }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
int check_problem(const magi_problem_t* pb) {
  if (!pb) return -1;
  int D, P;
  if (magi_b200_model_dims(pb->model_id, &D, &P) != 0) return MAGI_ERR_UNSUPPORTED;
  if (pb->D != D || pb->P != P) return -1;
  if (pb->B <= 0 || pb->R <= 0 || pb->n <= 1) return -1;
  if (!pb->packed || !pb->mu || !pb->y || !pb->mask || !pb->N_ds || !pb->beta || !pb->LB) return -1;
  if (pb->B > 65535) return MAGI_ERR_UNSUPPORTED;  // gridDim.y
  return MAGI_OK;
}

// ---- fast path dispatch (posterior_fast.cuh) ----
template <class M>
constexpr bool fast_model() { return M::D <= kFastMaxD; }
template <class M>
bool use_fast(int n) { return fast_model<M>() && magi_pad8(n) <= kFastMaxNp; }

int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}
// persistent grid: one CTA per SM (or fewer when there are fewer items)
int fast_grid(const magi_problem_t* pb) {
  const long items = (long)pb->B * ((pb->R + kCh - 1) / kCh);
  const int sms = sm_count();
  return (int)(items < sms ? items : sms);
}

template <class M>
size_t workspace_bytes_t(const magi_problem_t* pb) {
  if (use_fast<M>(pb->n)) return (size_t)sm_count() * 3 * fast_slot_elems<M>(magi_pad8(pb->n)) * sizeof(double);
  const Placement<M> p = placement<M>(pb->n, true, true);
  const size_t ncta = (size_t)pb->B * ((pb->R + kCh - 1) / kCh);
  return ncta * (p.ws_big_elems + p.ws_save_elems) * sizeof(double);
}

template <class M>
int launch_logpost(const magi_problem_t* pb, const double* X, const double* sig_pre, const double* th_pre,
                   const double* beta_temp, double* lp, double* gX, double* gsig, double* gth, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  if constexpr (fast_model<M>()) {
    if (use_fast<M>(pb->n)) {
      const int np = magi_pad8(pb->n);
      const size_t smem = FastScratch<M, 0>::elems(np, pb->band) * sizeof(double);
      auto kern = np == kFastMaxNp ? logpost_grad_fast_kernel<M, kFastMaxNp> : logpost_grad_fast_kernel<M, 0>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return magi_cuda_status(e);
      kern<<<fast_grid(pb), 32 * (np / 8 + kTsProd), smem, st>>>(*pb, X, sig_pre, th_pre, beta_temp, lp, gX, gsig, gth);
      return magi_cuda_status(cudaGetLastError());
    }
  }
  const Placement<M> p = placement<M>(pb->n, false, false);
  const dim3 grid((pb->R + kCh - 1) / kCh, pb->B);
  if (p.ws_big_elems && (!ws || ws_bytes < (size_t)grid.x * grid.y * p.ws_big_elems * sizeof(double))) return -10;
  auto kern = p.big_in_smem ? logpost_grad_kernel<M, true> : logpost_grad_kernel<M, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
  if (e != cudaSuccess) return magi_cuda_status(e);
#ifdef MAGI_CARVEOUT
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, MAGI_CARVEOUT);
#endif
  kern<<<grid, threads_for(pb->n), p.smem_bytes, st>>>(*pb, X, sig_pre, th_pre, beta_temp, lp, gX, gsig, gth,
                                                       static_cast<double*>(ws), p.ws_big_elems);
  return magi_cuda_status(cudaGetLastError());
}

template <class M>
int launch_leapfrog(const magi_problem_t* pb, double* X, double* sig_pre, double* th_pre, double* pX,
                    double* psig, double* pth, const double* eps, const double* beta_temp, int n_steps,
                    double* lp_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if constexpr (fast_model<M>()) {
    if (use_fast<M>(pb->n)) {
      const int np = magi_pad8(pb->n);
      const size_t smem = FastScratch<M, 0>::elems(np, pb->band) * sizeof(double);
      const int grid = fast_grid(pb);
      if (!ws || ws_bytes < (size_t)grid * fast_slot_elems<M>(np) * sizeof(double)) return -12;
      auto kern = np == kFastMaxNp ? leapfrog_fast_kernel<M, kFastMaxNp> : leapfrog_fast_kernel<M, 0>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return magi_cuda_status(e);
      kern<<<grid, 32 * (np / 8 + kTsProd), smem, st>>>(*pb, X, sig_pre, th_pre, pX, psig, pth, eps, beta_temp, n_steps, lp_out,
                                             static_cast<double*>(ws));
      return magi_cuda_status(cudaGetLastError());
    }
  }
  const Placement<M> p = placement<M>(pb->n, true, false);
  const dim3 grid((pb->R + kCh - 1) / kCh, pb->B);
  if (p.ws_big_elems && (!ws || ws_bytes < (size_t)grid.x * grid.y * p.ws_big_elems * sizeof(double))) return -12;
  auto kern = p.big_in_smem ? leapfrog_kernel<M, true> : leapfrog_kernel<M, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
  if (e != cudaSuccess) return magi_cuda_status(e);
  kern<<<grid, threads_for(pb->n), p.smem_bytes, st>>>(*pb, X, sig_pre, th_pre, pX, psig, pth, eps, beta_temp, n_steps,
                                                       lp_out, static_cast<double*>(ws), p.ws_big_elems);
  return magi_cuda_status(cudaGetLastError());
}

template <class M>
int launch_hmc(const magi_problem_t* pb, const magi_hmc_config_t* cfg, double* X, double* sig_pre,
               double* th_pre, double* eps, double* da_state, HmcOut out, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  if constexpr (fast_model<M>()) {
    if (use_fast<M>(pb->n)) {
      const int np = magi_pad8(pb->n);
      const size_t smem = FastScratch<M, 0>::elems(np, pb->band) * sizeof(double);
      const int grid = fast_grid(pb);
      if (!ws || ws_bytes < (size_t)grid * 3 * fast_slot_elems<M>(np) * sizeof(double)) return -15;
      auto kern = np == kFastMaxNp ? hmc_fast_kernel<M, kFastMaxNp> : hmc_fast_kernel<M, 0>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return magi_cuda_status(e);
      kern<<<grid, 32 * (np / 8 + kTsProd), smem, st>>>(*pb, *cfg, X, sig_pre, th_pre, eps, da_state, out,
                                             static_cast<double*>(ws));
      return magi_cuda_status(cudaGetLastError());
    }
  }
  const Placement<M> p = placement<M>(pb->n, true, true);
  const dim3 grid((pb->R + kCh - 1) / kCh, pb->B);
  const size_t need = (size_t)grid.x * grid.y * (p.ws_big_elems + p.ws_save_elems) * sizeof(double);
  if (!ws || ws_bytes < need) return -15;
  auto kern = p.big_in_smem ? hmc_kernel<M, true> : hmc_kernel<M, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
  if (e != cudaSuccess) return magi_cuda_status(e);
  kern<<<grid, threads_for(pb->n), p.smem_bytes, st>>>(*pb, *cfg, X, sig_pre, th_pre, eps, da_state, out,
                                                       static_cast<double*>(ws), p.ws_big_elems, p.ws_save_elems);
  return magi_cuda_status(cudaGetLastError());
}

#ifdef MAGI_DEV_SEIR4_ONLY   // development builds (tools/build_variant.sh): one model, compiles in a fraction of the time
#define MAGI_DISPATCH_MODEL(id, EXPR)                  \
  switch (id) {                                        \
    case MAGI_MODEL_SEIR4: { using M = Seir4; EXPR; }  \
    default: return MAGI_ERR_UNSUPPORTED;              \
  }
#else
#define MAGI_DISPATCH_MODEL(id, EXPR)                  \
  switch (id) {                                        \
    case MAGI_MODEL_SEIR3: { using M = Seir3; EXPR; }  \
    case MAGI_MODEL_SEIR4: { using M = Seir4; EXPR; }  \
    case MAGI_MODEL_SIRW: { using M = Sirw; EXPR; }    \
    case MAGI_MODEL_LORENZ96: { using M = Lorenz96; EXPR; } \
    default: return MAGI_ERR_UNSUPPORTED;              \
  }
#endif

}  // namespace

extern "C" {

int magi_b200_abi_version(void) { return MAGI_B200_ABI_VERSION; }

#ifdef MAGI_TRACE
// instrumented builds only (tools/trace_fast.py): fetch and reset the event timeline of CTA 0
__attribute__((visibility("default"))) int magi_b200_debug_trace(unsigned long long* host_events, int* host_counts) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_events, g_magi_trace, sizeof(unsigned long long) * kFastMaxBlk0 * kTraceCap);
  cudaMemcpyFromSymbol(host_counts, g_magi_trace_n, sizeof(int) * kFastMaxBlk0);
  int zero[kFastMaxBlk0] = {0};
  cudaMemcpyToSymbol(g_magi_trace_n, zero, sizeof(zero));
  return (int)cudaDeviceSynchronize();
}
#endif

int magi_b200_model_dims(int model_id, int* D, int* P) {
  int d, p;
  switch (model_id) {
    case MAGI_MODEL_SEIR3: d = Seir3::D; p = Seir3::P; break;
    case MAGI_MODEL_SEIR4: d = Seir4::D; p = Seir4::P; break;
    case MAGI_MODEL_SIRW: d = Sirw::D; p = Sirw::P; break;
    case MAGI_MODEL_LORENZ96: d = Lorenz96::D; p = Lorenz96::P; break;
    default: return -1;
  }
  if (D) *D = d;
  if (P) *P = p;
  return 0;
}

const char* magi_b200_status_string(int status) {
  if (status == MAGI_OK) return "ok";
  if (status < 0) return "invalid argument (index = -status)";
  if (status == MAGI_ERR_UNSUPPORTED) return "unsupported model or shape";
  if (status >= MAGI_ERR_CUDA) return cudaGetErrorString((cudaError_t)(status - MAGI_ERR_CUDA));
  return "unknown status";
}

size_t magi_b200_packed_bytes(int B, int D, int n) {
  return (size_t)B * D * 3 * magi_packed_mat_elems(n) * sizeof(double);
}

int magi_b200_pack_matrices(const double* Cinv, const double* m, const double* Kinv, int B, int D, int n,
                            void* packed, magi_stream_t stream) {
  if (!Cinv) return -1;
  if (!m) return -2;
  if (!Kinv) return -3;
  if (B <= 0) return -4;
  if (D <= 0) return -5;
  if (n <= 1) return -6;
  if (!packed) return -7;
  const int np = magi_pad8(n);
  const size_t nmat = (size_t)B * D;
  for (size_t off = 0; off < nmat; off += 65535) {
    const unsigned nz = (unsigned)((nmat - off) < 65535 ? (nmat - off) : 65535);
    const dim3 grid((np + 255) / 256, np, nz);
    pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        Cinv + off * (size_t)n * n, m + off * (size_t)n * n, Kinv + off * (size_t)n * n, n, np,
        static_cast<double*>(packed) + off * 3 * (size_t)np * np);
  }
  return magi_cuda_status(cudaGetLastError());
}

size_t magi_b200_sampler_workspace_bytes(const magi_problem_t* pb) {
  int D, P;
  if (!pb || magi_b200_model_dims(pb->model_id, &D, &P) != 0) return 0;
  MAGI_DISPATCH_MODEL(pb->model_id, return workspace_bytes_t<M>(pb));
}

int magi_b200_logpost_grad(const magi_problem_t* pb, const double* X, const double* sig_pre,
                           const double* th_pre, const double* beta_temp, double* lp, double* gX, double* gsig,
                           double* gth, void* ws, size_t ws_bytes, magi_stream_t stream) {
  const int c = check_problem(pb);
  if (c != MAGI_OK) return c;
  if (!X) return -2;
  if (!sig_pre) return -3;
  if (!th_pre) return -4;
  if (!beta_temp) return -5;
  if (!lp) return -6;
  if (!gX) return -7;
  if (!gsig) return -8;
  if (!gth) return -9;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAGI_DISPATCH_MODEL(pb->model_id,
                      return launch_logpost<M>(pb, X, sig_pre, th_pre, beta_temp, lp, gX, gsig, gth, ws, ws_bytes, st));
}

int magi_b200_leapfrog(const magi_problem_t* pb, double* X, double* sig_pre, double* th_pre, double* pX,
                       double* psig, double* pth, const double* eps, const double* beta_temp, int n_steps,
                       double* lp_out, void* ws, size_t ws_bytes, magi_stream_t stream) {
  const int c = check_problem(pb);
  if (c != MAGI_OK) return c;
  if (!X) return -2;
  if (!sig_pre) return -3;
  if (!th_pre) return -4;
  if (!pX) return -5;
  if (!psig) return -6;
  if (!pth) return -7;
  if (!eps) return -8;
  if (!beta_temp) return -9;
  if (n_steps < 0) return -10;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAGI_DISPATCH_MODEL(pb->model_id, return launch_leapfrog<M>(pb, X, sig_pre, th_pre, pX, psig, pth, eps, beta_temp,
                                                              n_steps, lp_out, ws, ws_bytes, st));
}

int magi_b200_hmc_run(const magi_problem_t* pb, const magi_hmc_config_t* cfg, double* X, double* sig_pre,
                      double* th_pre, double* eps, double* da_state, double* th_samps, double* sig_samps,
                      double* X_samps, double* X_sum, double* X_sumsq, double* accept_prob, double* lp_trace,
                      void* ws, size_t ws_bytes, magi_stream_t stream) {
  const int c = check_problem(pb);
  if (c != MAGI_OK) return c;
  if (!cfg || cfg->n_iter < 0 || cfg->n_leapfrog < 1) return -2;
  if (!X) return -3;
  if (!sig_pre) return -4;
  if (!th_pre) return -5;
  if (!eps) return -6;
  if (!da_state) return -7;
  HmcOut out{th_samps, sig_samps, X_samps, X_sum, X_sumsq, accept_prob, lp_trace};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MAGI_DISPATCH_MODEL(pb->model_id,
                      return launch_hmc<M>(pb, cfg, X, sig_pre, th_pre, eps, da_state, out, ws, ws_bytes, st));
}

}  // extern "C"
