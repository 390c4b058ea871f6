// Fused per-leaf bookkeeping of the batched No-U-Turn sampler (include/magi_b200_nuts.h).
// Pure streaming kernels: every [C,S] array is touched once per leaf, the per-chain scalars (energy, weights,
// U-turn dot products) are block reductions inside the same pass.  HBM-bound: ~5 array passes in `pre`, ~9-11 in
// `post` (4 reads + 4 writes + 2 per checkpoint touched) of C*S*8 bytes each.
#include "common.cuh"
#include "rng.cuh"
#include "../../include/magi_b200_nuts.h"

namespace {

constexpr int kThreads = 256;

struct Parts {
  const double* X;
  const double* s;
  const double* t;
};

__device__ __forceinline__ double part_load(const Parts& p, int c, int i, int nD, int D, int P) {
  if (i < nD) return p.X[(size_t)c * nD + i];
  if (i < nD + D) return p.s[(size_t)c * D + (i - nD)];
  return p.t[(size_t)c * P + (i - nD - D)];
}

__global__ void __launch_bounds__(kThreads) nuts_leaf_pre_kernel(magi_nuts_subtree_t st, double* __restrict__ ph,
                                                                 double* __restrict__ Xn, double* __restrict__ sn,
                                                                 double* __restrict__ tn) {
  const int c = blockIdx.x;
  const int nD = st.nD, D = st.D, P = st.P, S = nD + D + P;
  const bool on = st.building[c] != 0;
  const double e = st.e[c];
  const double* zc = st.zc + (size_t)c * S;
  const double* pc = st.pc + (size_t)c * S;
  const double* gc = st.gc + (size_t)c * S;
  double* phc = ph + (size_t)c * S;
  for (int i = threadIdx.x; i < S; i += kThreads) {
    double z = zc[i];
    if (on) {
      const double h = fma(0.5 * e, gc[i], pc[i]);
      phc[i] = h;
      z = fma(e, h, z);
    }
    if (i < nD) Xn[(size_t)c * nD + i] = z;
    else if (i < nD + D) sn[(size_t)c * D + (i - nD)] = z;
    else tn[(size_t)c * P + (i - nD - D)] = z;
  }
}

struct Checks {
  int n;
  int slot[MAGI_NUTS_MAX_CHECKS];
};

__device__ __forceinline__ double logaddexp_d(double a, double b) {
  const double m = fmax(a, b);
  if (!(m > -INFINITY)) return m;  // both -inf (or NaN)
  return m + log1p(exp(-fabs(a - b)));
}

struct PartsW {
  double* X;
  double* s;
  double* t;
};

__device__ __forceinline__ void part_store(const PartsW& p, int c, int i, int nD, int D, int P, double v) {
  if (i < nD) p.X[(size_t)c * nD + i] = v;
  else if (i < nD + D) p.s[(size_t)c * D + (i - nD)] = v;
  else p.t[(size_t)c * P + (i - nD - D)] = v;
}

// `next` != 0: the first half of the NEXT leaf's leapfrog step is done here as well (ph <- p + e/2 g, the position
// parts <- z + e ph), so that only the first leaf of a subtree needs `nuts_leaf_pre_kernel`.
template <int NCHK>
__global__ void __launch_bounds__(kThreads, NCHK <= 2 ? 4 : 2) nuts_leaf_post_kernel(magi_nuts_subtree_t st,
                                                                  double* __restrict__ ph, PartsW zn, Parts gn, int next,
                                                                  const double* __restrict__ lp_new,
                                                                  const double* __restrict__ log_u, int64_t log_u_stride,
                                                                  double max_energy_diff, int slot_store, Checks ck) {
  const int c = blockIdx.x;
  if (st.building[c] == 0) return;
  const int nD = st.nD, D = st.D, P = st.P, S = nD + D + P, C = st.C;
  const double he = 0.5 * st.e[c];
  const size_t row = (size_t)c * S;
  double* zc = st.zc + row;
  double* pc = st.pc + row;
  double* gc = st.gc + row;
  double* rho = st.rho_sub + row;
  double* phc = ph + row;
  const double e = st.e[c];
  double* store_p = slot_store >= 0 ? st.ck_p + ((size_t)slot_store * C + c) * S : nullptr;
  double* store_r = slot_store >= 0 ? st.ck_rho + ((size_t)slot_store * C + c) * S : nullptr;

  // acc[0] = p.p ; acc[1 + 2k] = rb_k . ck_p_k ; acc[2 + 2k] = rb_k . p
  double acc[1 + 2 * NCHK];
#pragma unroll
  for (int q = 0; q < 1 + 2 * NCHK; ++q) acc[q] = 0.0;

  for (int i = threadIdx.x; i < S; i += kThreads) {
    const double g = part_load(gn, c, i, nD, D, P);
    const double z = part_load(Parts{zn.X, zn.s, zn.t}, c, i, nD, D, P);
    const double p = fma(he, g, phc[i]);
    if (next) {
      const double h = fma(he, g, p);
      phc[i] = h;
      part_store(zn, c, i, nD, D, P, fma(e, h, z));
    }
    const double r_old = rho[i];
    const double r_new = r_old + p;
    zc[i] = z;
    pc[i] = p;
    gc[i] = g;
    rho[i] = r_new;
    if (store_p) {
      store_r[i] = r_old;
      store_p[i] = p;
    }
    acc[0] = fma(p, p, acc[0]);
#pragma unroll
    for (int k = 0; k < NCHK; ++k) {
      if (k < ck.n) {
        const size_t o = ((size_t)ck.slot[k] * C + c) * S + i;
        const double rb = r_new - st.ck_rho[o];
        acc[1 + 2 * k] = fma(rb, st.ck_p[o], acc[1 + 2 * k]);
        acc[2 + 2 * k] = fma(rb, p, acc[2 + 2 * k]);
      }
    }
  }

  __shared__ double red[kThreads / 32][1 + 2 * NCHK];
  __shared__ int s_take;
  const int nred = 1 + 2 * ck.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 1 + 2 * NCHK; ++q) {
    if (q < nred) {
      const double v = magi_warp_sum(acc[q]);
      if (lane == 0) red[warp][q] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[1 + 2 * NCHK];
    for (int q = 0; q < nred; ++q) {
      double v = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) v += red[w][q];
      tot[q] = v;
    }
    const double lp = lp_new[c];
    double dE = -lp + 0.5 * tot[0] - st.H0[c];
    if (!isfinite(dE)) dE = INFINITY;
    st.sum_acc[c] += exp(fmin(-dE, 0.0));
    st.n_leaf[c] += 1;
    const bool div = dE > max_energy_diff;
    if (div) st.diverged[c] = 1;
    const double lw_new = logaddexp_d(st.logw_sub[c], -dE);
    const bool take = log_u[(size_t)c * log_u_stride] < (-dE - lw_new);   // false for NaN (-inf - -inf)
    st.logw_sub[c] = lw_new;
    if (take) st.sub_lp[c] = lp;
    bool ok = !div;
    for (int k = 0; k < ck.n; ++k) ok = ok && (tot[1 + 2 * k] > 0.0) && (tot[2 + 2 * k] > 0.0);
    st.building[c] = ok ? 1 : 0;
    s_take = take ? 1 : 0;
  }
  __syncthreads();
  if (s_take) {
    double* sz = st.sub_z + row;
    for (int i = threadIdx.x; i < S; i += kThreads) sz[i] = zc[i];
  }
}

__global__ void __launch_bounds__(kThreads) nuts_uniform_kernel(uint64_t seed, const int64_t* __restrict__ chain_ids,
                                                                uint32_t iteration, uint32_t purpose, uint32_t index0,
                                                                int count, int C, double* __restrict__ ua,
                                                                double* __restrict__ ub) {
  const size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= (size_t)C * count) return;
  const int c = (int)(t / count), k = (int)(t % count);
  const uint4 r = magi_philox(make_uint4(index0 + (uint32_t)k, (uint32_t)chain_ids[c], iteration, purpose),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  ua[t] = magi_u53(r.x, r.y);
  ub[t] = magi_u53(r.z, r.w);
}

__global__ void __launch_bounds__(kThreads) nuts_momentum_kernel(uint64_t seed, const int64_t* __restrict__ chain_ids,
                                                                 uint32_t iteration, int S, double* __restrict__ p0) {
  const int c = blockIdx.y;
  const int pair = blockIdx.x * kThreads + threadIdx.x;
  if (2 * pair >= S) return;
  double z0, z1;
  magi_normal_pair(seed, (uint32_t)pair, (uint32_t)chain_ids[c], iteration, z0, z1);
  double* row = p0 + (size_t)c * S;
  row[2 * pair] = z0;
  if (2 * pair + 1 < S) row[2 * pair + 1] = z1;
}

__global__ void __launch_bounds__(kThreads) nuts_begin_kernel(magi_nuts_subtree_t st, magi_nuts_tree_t tr) {
  const int c = blockIdx.x;
  const int S = st.nD + st.D + st.P;
  const size_t row = (size_t)c * S;
  const bool fwd = tr.fwd[c] != 0;
  const double* sz = (fwd ? tr.zr : tr.zl) + row;
  const double* sp = (fwd ? tr.pr : tr.pl) + row;
  const double* sg = (fwd ? tr.gr : tr.gl) + row;
  for (int i = threadIdx.x; i < S; i += kThreads) {
    const double z = sz[i];
    st.zc[row + i] = z;
    st.pc[row + i] = sp[i];
    st.gc[row + i] = sg[i];
    st.rho_sub[row + i] = 0.0;
    st.sub_z[row + i] = z;
  }
}

__global__ void __launch_bounds__(kThreads) nuts_merge_kernel(magi_nuts_subtree_t st, magi_nuts_tree_t tr,
                                                              const double* __restrict__ log_u_acc) {
  const int c = blockIdx.x;
  if (st.building[c] == 0) {
    if (threadIdx.x == 0) tr.active[c] = 0;
    return;
  }
  const int S = st.nD + st.D + st.P;
  const size_t row = (size_t)c * S;
  const bool fwd = tr.fwd[c] != 0;
  const double lw_sub = st.logw_sub[c], lw = tr.logw[c];
  const bool swap = log_u_acc[c] < (lw_sub - lw);
  double* dz = (fwd ? tr.zr : tr.zl) + row;
  double* dp = (fwd ? tr.pr : tr.pl) + row;
  double* dg = (fwd ? tr.gr : tr.gl) + row;
  const double* op = (fwd ? tr.pl : tr.pr) + row;
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < S; i += kThreads) {
    const double r = tr.rho[row + i] + st.rho_sub[row + i];
    const double p = st.pc[row + i];
    tr.rho[row + i] = r;
    dz[i] = st.zc[row + i];
    dp[i] = p;
    dg[i] = st.gc[row + i];
    if (swap) tr.prop_z[row + i] = st.sub_z[row + i];
    a0 = fma(r, p, a0);
    a1 = fma(r, op[i], a1);
  }
  __shared__ double red[kThreads / 32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a0 = magi_warp_sum(a0);
  a1 = magi_warp_sum(a1);
  if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { t0 += red[w][0]; t1 += red[w][1]; }
    tr.logw[c] = logaddexp_d(lw, lw_sub);
    if (swap) tr.prop_lp[c] = st.sub_lp[c];
    tr.active[c] = (t0 > 0.0 && t1 > 0.0) ? 1 : 0;
  }
}

int check_tree(const magi_nuts_tree_t* t) {
  if (!t) return -2;
  if (!t->zl || !t->pl || !t->gl || !t->zr || !t->pr || !t->gr || !t->rho || !t->prop_z || !t->prop_lp || !t->logw ||
      !t->active || !t->fwd)
    return -2;
  return MAGI_OK;
}

int check_subtree(const magi_nuts_subtree_t* st) {
  if (!st) return -1;
  if (st->C <= 0 || st->nD <= 0 || st->D <= 0 || st->P < 0) return -1;
  if (!st->zc || !st->pc || !st->gc || !st->rho_sub || !st->sub_z || !st->sub_lp || !st->logw_sub || !st->sum_acc ||
      !st->n_leaf || !st->building || !st->diverged || !st->e || !st->H0)
    return -1;
  return MAGI_OK;
}

}  // namespace

extern "C" int magi_b200_nuts_leaf_pre(const magi_nuts_subtree_t* st, double* ph, double* Xn, double* sn, double* tn,
                                       magi_stream_t stream) {
  if (int s = check_subtree(st)) return s;
  if (!ph) return -2;
  if (!Xn) return -3;
  if (!sn) return -4;
  if (!tn && st->P > 0) return -5;
  nuts_leaf_pre_kernel<<<st->C, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, ph, Xn, sn, tn);
  return magi_cuda_status(cudaGetLastError());
}

extern "C" int magi_b200_nuts_leaf_post(const magi_nuts_subtree_t* st, const double* ph, const double* Xn,
                                        const double* sn, const double* tn, const double* lp_new, const double* gX,
                                        const double* gs, const double* gt, const double* log_u, int64_t log_u_stride,
                                        double max_energy_diff, int slot_store, int n_checks, const int* check_slots,
                                        magi_stream_t stream) {
  return magi_b200_nuts_leaf_post_next(st, const_cast<double*>(ph), const_cast<double*>(Xn), const_cast<double*>(sn),
                                       const_cast<double*>(tn), lp_new, gX, gs, gt, log_u, log_u_stride, max_energy_diff,
                                       slot_store, n_checks, check_slots, 0, stream);
}

extern "C" int magi_b200_nuts_leaf_post_next(const magi_nuts_subtree_t* st, double* ph, double* Xn, double* sn,
                                             double* tn, const double* lp_new, const double* gX, const double* gs,
                                             const double* gt, const double* log_u, int64_t log_u_stride,
                                             double max_energy_diff, int slot_store, int n_checks,
                                             const int* check_slots, int next, magi_stream_t stream) {
  if (int s = check_subtree(st)) return s;
  if (!ph) return -2;
  if (!Xn) return -3;
  if (!sn) return -4;
  if (!tn && st->P > 0) return -5;
  if (!lp_new) return -6;
  if (!gX) return -7;
  if (!gs) return -8;
  if (!gt && st->P > 0) return -9;
  if (!log_u) return -10;
  if (n_checks < 0 || n_checks > MAGI_NUTS_MAX_CHECKS) return -14;
  if ((slot_store >= 0 || n_checks > 0) && (!st->ck_p || !st->ck_rho)) return -1;
  if (n_checks > 0 && !check_slots) return -15;
  Checks ck;
  ck.n = n_checks;
  for (int k = 0; k < MAGI_NUTS_MAX_CHECKS; ++k) ck.slot[k] = k < n_checks ? check_slots[k] : 0;
  const PartsW zn{Xn, sn, tn};
  const Parts gn{gX, gs, gt};
  const cudaStream_t cs = static_cast<cudaStream_t>(stream);
#define MAGI_POST(N) \
  nuts_leaf_post_kernel<N><<<st->C, kThreads, 0, cs>>>(*st, ph, zn, gn, next, lp_new, log_u, log_u_stride, max_energy_diff, \
                                                      slot_store, ck)
  if (n_checks == 0) MAGI_POST(0);
  else if (n_checks == 1) MAGI_POST(1);
  else if (n_checks == 2) MAGI_POST(2);
  else if (n_checks <= 4) MAGI_POST(4);
  else MAGI_POST(MAGI_NUTS_MAX_CHECKS);
#undef MAGI_POST
  return magi_cuda_status(cudaGetLastError());
}

extern "C" int magi_b200_nuts_momentum(uint64_t seed, const int64_t* chain_ids, uint32_t iteration, int C, int S,
                                       double* p0, magi_stream_t stream) {
  if (!chain_ids) return -2;
  if (C <= 0 || C > 65535 * 32) return -4;
  if (S <= 0) return -5;
  if (!p0) return -6;
  const int npair = (S + 1) / 2;
  // gridDim.y <= 65535: launch in slabs of chains
  for (int c0 = 0; c0 < C; c0 += 65535) {
    const int nc = C - c0 < 65535 ? C - c0 : 65535;
    const dim3 grid((npair + kThreads - 1) / kThreads, nc);
    nuts_momentum_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(seed, chain_ids + c0, iteration, S,
                                                                                  p0 + (size_t)c0 * S);
  }
  return magi_cuda_status(cudaGetLastError());
}

namespace {
// one CTA per chain: p += kick eps g (, z += eps p) (, energy = 1/2 |p|^2)
__global__ void __launch_bounds__(kThreads) hmc_kick_drift_kernel(int S, double* __restrict__ z, double* __restrict__ p,
                                                                 const double* __restrict__ g,
                                                                 const double* __restrict__ eps, double kick, int drift,
                                                                 double* __restrict__ energy) {
  const size_t c = blockIdx.x;
  const double e = eps[c], h = kick * e;
  double ke = 0.0;
  for (int i = threadIdx.x; i < S; i += kThreads) {
    const size_t o = c * S + i;
    const double pn = fma(h, g[o], p[o]);
    p[o] = pn;
    if (drift) z[o] = fma(e, pn, z[o]);
    ke = fma(pn, pn, ke);
  }
  if (energy) {
    __shared__ double part[kThreads / 32];
    ke = magi_warp_sum(ke);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ke;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) t += part[w];
      energy[c] = 0.5 * t;
    }
  }
}
}  // namespace

extern "C" int magi_b200_hmc_kick_drift(int C, int S, double* z, double* p, const double* g, const double* eps, double kick,
                                        int drift, double* energy, magi_stream_t stream) {
  if (C <= 0) return -1;
  if (S <= 0) return -2;
  if (!z && drift) return -3;
  if (!p) return -4;
  if (!g) return -5;
  if (!eps) return -6;
  hmc_kick_drift_kernel<<<C, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(S, z, p, g, eps, kick, drift, energy);
  return magi_cuda_status(cudaGetLastError());
}

extern "C" int magi_b200_nuts_subtree_begin(const magi_nuts_subtree_t* st, const magi_nuts_tree_t* tree,
                                            magi_stream_t stream) {
  if (int s = check_subtree(st)) return s;
  if (int s = check_tree(tree)) return s;
  nuts_begin_kernel<<<st->C, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, *tree);
  return magi_cuda_status(cudaGetLastError());
}

extern "C" int magi_b200_nuts_merge(const magi_nuts_subtree_t* st, const magi_nuts_tree_t* tree, const double* log_u_acc,
                                    magi_stream_t stream) {
  if (int s = check_subtree(st)) return s;
  if (int s = check_tree(tree)) return s;
  if (!log_u_acc) return -3;
  nuts_merge_kernel<<<st->C, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, *tree, log_u_acc);
  return magi_cuda_status(cudaGetLastError());
}

extern "C" int magi_b200_nuts_uniforms(uint64_t seed, const int64_t* chain_ids, uint32_t iteration, uint32_t purpose,
                                       uint32_t index0, int count, int C, double* ua, double* ub,
                                       magi_stream_t stream) {
  if (!chain_ids) return -2;
  if (count <= 0) return -6;
  if (C <= 0) return -7;
  if (!ua) return -8;
  if (!ub) return -9;
  const size_t total = (size_t)C * count;
  nuts_uniform_kernel<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      seed, chain_ids, iteration, purpose, index0, count, C, ua, ub);
  return magi_cuda_status(cudaGetLastError());
}
