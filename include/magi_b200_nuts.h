/*
 * magi_b200_nuts.h -- C ABI of the fused per-leaf bookkeeping of the batched No-U-Turn sampler.
 *
 * The reference samples with tfp.mcmc.NoUTurnSampler (magi_v2.py:360-366, :866-869; tensorflow-probability 0.24.0,
 * mcmc/nuts.py -- third party, not under /root/reference).  TFP's batched NUTS advances every chain by one leapfrog
 * step per iteration of its while-loop and does the tree bookkeeping with masked tensor ops; here the same loop body
 * is three launches per leaf: `magi_b200_nuts_leaf_pre` (first half of the leapfrog step, writes the new position in
 * the layout `magi_b200_logpost_grad` reads), `magi_b200_logpost_grad` (magi_b200.h), `magi_b200_nuts_leaf_post`
 * (second half of the step + everything TFP's loop body does per leaf).  Conventions as in magi_b200.h: device
 * pointers, binary64, caller-owned buffers, asynchronous on `stream`, status return, no CPU fallback.
 *
 * State layout: a chain's state is z = [X (n*D, time-major) | sigma_sqs_pre (D) | thetas_pre (P)], S = n*D + D + P
 * doubles (magi_v2.py:383); "packed" arrays are [C, S] row-major over C = B*R chains; "parts" are the three separate
 * contiguous arrays [C, n*D], [C, D], [C, P] that magi_b200_logpost_grad takes and returns.
 */
#ifndef MAGI_B200_NUTS_H
#define MAGI_B200_NUTS_H

#include "magi_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MAGI_NUTS_MAX_CHECKS 12

/* The running state of the subtree being built (one leaf per call), all [C] or [C, S] device arrays. */
typedef struct {
  int C, nD, D, P;        /* chains and the split of S = nD + D + P */
  double* zc;             /* [C,S] in/out: position, momentum and gradient at the subtree's newest leaf */
  double* pc;
  double* gc;
  double* rho_sub;        /* [C,S] in/out: sum of the momenta of the subtree's leaves */
  double* sub_z;          /* [C,S] in/out: the subtree's multinomial proposal ... */
  double* sub_lp;         /* [C]   ... and its log-posterior */
  double* logw_sub;       /* [C]   in/out: log of the subtree's total weight sum_i exp(H0 - H_i) */
  double* sum_acc;        /* [C]   in/out: sum over leaves of min(1, exp(H0 - H)) (dual-averaging statistic) */
  int64_t* n_leaf;        /* [C]   in/out: leaves evaluated */
  uint8_t* building;      /* [C]   in/out: 1 while the chain's subtree is valid and still being built */
  uint8_t* diverged;      /* [C]   in/out: H - H0 exceeded max_energy_diff somewhere in this transition */
  double* ck_p;           /* [n_slots, C, S] checkpoint memory: momentum of the even leaf that opened slot k ... */
  double* ck_rho;         /* [n_slots, C, S] ... and the subtree's momentum sum before that leaf */
  const double* e;        /* [C] signed step size (direction * eps) */
  const double* H0;       /* [C] energy at the start of the transition */
} magi_nuts_subtree_t;

/* First half of one leapfrog step for every chain that is still building (TFP SimpleLeapfrogIntegrator, identity
 * mass): ph = pc + e/2 * gc (packed [C,S]); z_new = zc + e * ph written as parts Xn [C,nD], sn [C,D], tn [C,P].
 * Chains that are not building still get their (unchanged) zc copied to the parts so that the evaluation that
 * follows stays finite. */
MAGI_API int magi_b200_nuts_leaf_pre(const magi_nuts_subtree_t* st, double* ph, double* Xn, double* sn, double* tn,
                                     magi_stream_t stream);

/* Second half of the step + the per-leaf bookkeeping, one CTA per chain; chains with building = 0 are skipped.
 *   p_new = ph + e/2 * g_new;  dE = -lp_new + p_new.p_new / 2 - H0  (non-finite -> +inf)
 *   sum_acc += min(1, exp(-dE)); n_leaf += 1; diverged |= dE > max_energy_diff
 *   the leaf replaces the subtree's proposal when log_u < -dE - logaddexp(logw_sub, -dE); logw_sub updated
 *   (zc, pc, gc) <- the new leaf; rho_sub += p_new
 *   leaf_index even: ck_rho[slot_store] <- rho_sub before this leaf, ck_p[slot_store] <- p_new
 *   leaf_index odd: for each of the n_checks dyadic blocks that end here (slots check_slots[k]):
 *       rb = rho_sub - ck_rho[slot];  turning |= !(rb . ck_p[slot] > 0 && rb . p_new > 0)
 *   building <- !diverging && !turning
 * lp_new [C], (gX, gs, gt) and (Xn, sn, tn) are the parts magi_b200_logpost_grad wrote / read; log_u [C] with element
 * stride log_u_stride. */
MAGI_API int magi_b200_nuts_leaf_post(const magi_nuts_subtree_t* st, const double* ph, const double* Xn,
                                      const double* sn, const double* tn, const double* lp_new, const double* gX,
                                      const double* gs, const double* gt, const double* log_u, int64_t log_u_stride,
                                      double max_energy_diff, int slot_store, int n_checks,
                                      const int* check_slots /* host, n_checks entries */, magi_stream_t stream);

/* The tree of the current transition, [C,S] / [C] device arrays. */
typedef struct {
  double* zl;             /* [C,S] leftmost state: position, momentum, gradient */
  double* pl;
  double* gl;
  double* zr;             /* [C,S] rightmost state */
  double* pr;
  double* gr;
  double* rho;            /* [C,S] sum of the momenta of every leaf of the tree */
  double* prop_z;         /* [C,S] the transition's current proposal ... */
  double* prop_lp;        /* [C]   ... and its log-posterior */
  double* logw;           /* [C]   log of the tree's total weight */
  uint8_t* active;        /* [C]   out of `merge`: 1 = the chain goes on doubling */
  const uint8_t* fwd;     /* [C]   direction of the current doubling: 1 = forward (extends the right end) */
} magi_nuts_tree_t;

/* Momentum draw of a transition: p0 [C,S] standard normals from the sampler's counter-based stream
 * (Philox4x32-10, counter = (pair index, chain_ids[c], iteration, 0), key = seed; Box-Muller) -- the same numbers
 * `magi_b200_hmc_run` draws for that chain and iteration.  chain_ids: [C] int64 device array. */
MAGI_API int magi_b200_nuts_momentum(uint64_t seed, const int64_t* chain_ids, uint32_t iteration, int C, int S,
                                     double* p0, magi_stream_t stream);

/* One fused step of TFP's SimpleLeapfrogIntegrator for the host-driven fixed-length HMC (magi_v2_b200/hmc_host.py:
 * grids too large for the fused sampler kernel):  p += kick * eps[c] * g;  if drift: z += eps[c] * p.
 * kick = 1/2 for the first and the last half step, 1 when the two half kicks of consecutive steps are applied as one.
 * z, p, g: [C,S]; eps: [C].  Also returns, if energy != NULL, energy[c] = 1/2 |p_c|^2 after the kick (drift == 0). */
MAGI_API int magi_b200_hmc_kick_drift(int C, int S, double* z, double* p, const double* g, const double* eps, double kick,
                                      int drift, double* energy, magi_stream_t stream);

/* `magi_b200_nuts_leaf_post` that, with next != 0, also performs the first half of the NEXT leaf's step for the chains
 * it updates (ph <- p_new + e/2 g_new; Xn, sn, tn <- z_new + e ph): then only the first leaf of a subtree needs
 * `magi_b200_nuts_leaf_pre`.  ph, Xn, sn, tn are read and written. */
MAGI_API int magi_b200_nuts_leaf_post_next(const magi_nuts_subtree_t* st, double* ph, double* Xn, double* sn, double* tn,
                                           const double* lp_new, const double* gX, const double* gs, const double* gt,
                                           const double* log_u, int64_t log_u_stride, double max_energy_diff,
                                           int slot_store, int n_checks, const int* check_slots /* host */, int next,
                                           magi_stream_t stream);

/* Uniform draws of the tree builder from the same stream: ua[c,k], ub[c,k] = the two (0,1) doubles of
 * Philox(counter = (index0 + k, chain_ids[c], iteration, purpose)), k < count.  purpose 2, index = doubling j: ua decides
 * the direction (ua < 1/2 = forward), ub the acceptance of the completed subtree; purpose 3, index = number of the leaf
 * within the transition: ua selects the proposal inside the subtree. */
MAGI_API int magi_b200_nuts_uniforms(uint64_t seed, const int64_t* chain_ids, uint32_t iteration, uint32_t purpose,
                                     uint32_t index0, int count, int C, double* ua, double* ub, magi_stream_t stream);

/* Start of a doubling: (zc, pc, gc) <- the end of the tree the direction points to, rho_sub <- 0, sub_z <- zc. */
MAGI_API int magi_b200_nuts_subtree_begin(const magi_nuts_subtree_t* st, const magi_nuts_tree_t* tree,
                                          magi_stream_t stream);

/* End of a doubling, for the chains whose subtree completed (building = 1; the others get active = 0 and keep their
 * tree): biased progressive sampling (the subtree's proposal replaces the tree's when log_u_acc < logw_sub - logw),
 * logw <- logaddexp(logw, logw_sub), rho += rho_sub, the extended end <- (zc, pc, gc), and the generalised U-turn
 * criterion of the whole tree: active <- rho.p_left > 0 && rho.p_right > 0.  log_u_acc [C]. */
MAGI_API int magi_b200_nuts_merge(const magi_nuts_subtree_t* st, const magi_nuts_tree_t* tree, const double* log_u_acc,
                                  magi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGI_B200_NUTS_H */
