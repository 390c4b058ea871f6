/*
 * magi_b200_wide.h -- the log-posterior + gradient of magi_b200.h for FEW datasets.
 *
 * `magi_b200_logpost_grad` gives one CTA a whole dataset (all components, all rows): the right shape when there are
 * at least as many (dataset, chain-group) pairs as SMs (BASELINE config 4).  With one dataset (the reference's own
 * use: magi_v2.py:286-425 on a single series), the 20 datasets of config 2, or the n = 1281 grid of config 5, that
 * leaves most of the GPU idle and bounds an evaluation by what ONE SM can stream.  The entry point below computes the
 * same function (magi_v2.py:308-348 and its gradient; same arguments, same results to rounding) with the matrix rows
 * of every component spread over the whole grid: three passes (C^-1 x and m x -> r; K^-1 r; m^T q + the point-wise
 * assembly) separated by kernel boundaries, plus a final per-chain reduction.  Conventions as in magi_b200.h.
 */
#ifndef MAGI_B200_WIDE_H
#define MAGI_B200_WIDE_H

#include "magi_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Bytes of caller-owned workspace for `magi_b200_logpost_grad_wide` on this problem (host-only). */
MAGI_API size_t magi_b200_logpost_grad_wide_workspace_bytes(const magi_problem_t* prob);

/* Same contract as magi_b200_logpost_grad (magi_b200.h): X [B,R,n,D], sig_pre [B,R,D], th_pre [B,R,P],
 * beta_temp [B,R] -> lp [B,R], gX [B,R,n,D], gsig [B,R,D], gth [B,R,P].  Four kernel launches on `stream`. */
MAGI_API int magi_b200_logpost_grad_wide(const magi_problem_t* prob, const double* X, const double* sig_pre,
                                         const double* th_pre, const double* beta_temp, double* lp, double* gX,
                                         double* gsig, double* gth, void* workspace, size_t workspace_bytes,
                                         magi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGI_B200_WIDE_H */
