/*
 * magi_b200.h -- C ABI of the B200-native MAGI posterior-evaluation path.
 *
 * The reference (sophiaxxiao/magi_v2, one Python file on TensorFlow-Probability) has no FFI:
 * the hot path sits behind two Python call signatures.  Each entry point
 * below names the reference lines it replaces.  A maintainer binds these with ctypes (see
 * INTEGRATION.md); `magi_v2_b200/ops.py` wraps them as `torch.ops.magi_b200.*`.
 *
 * Conventions (all entry points):
 *   - every array pointer is a DEVICE pointer unless marked "host"; all reals are IEEE binary64,
 *     row-major, dense, contiguous; leading batch dims B = datasets, R = chains per dataset;
 *   - the caller owns every buffer, including workspaces (query the *_bytes functions);
 *     the library never allocates or frees device memory;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t); no internal synchronisation;
 *     no global mutable state: safe from several host threads on different streams/devices;
 *   - the return value is a status: 0 = ok, -k = argument k (1-based) is invalid,
 *     MAGI_ERR_CUDA (1000 + cudaError_t) = a CUDA runtime error at launch,
 *     MAGI_ERR_UNSUPPORTED = shape/model combination this build does not handle;
 *   - there is NO CPU fallback: without a CUDA device these calls return MAGI_ERR_CUDA + code.
 */
#ifndef MAGI_B200_H
#define MAGI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGI_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define MAGI_API __attribute__((visibility("default")))
#else
#define MAGI_API
#endif

#define MAGI_OK 0
#define MAGI_ERR_UNSUPPORTED 900
#define MAGI_ERR_CUDA 1000 /* + cudaError_t */

/* ODE right-hand sides compiled into the library: f, d f/d x and d f/d theta as device code.
 * They replace the user's Python/TF callable `f_vec(t, X, thetas)` (magi_v2.py:32-33, :73, :335)
 * and the reverse-mode autodiff TFP applies to it. */
typedef enum {
  MAGI_MODEL_SEIR3 = 0,    /* vignette.ipynb:68-79   (E,I,R; S implicit), theta=(beta,gamma,sigma) */
  MAGI_MODEL_SEIR4 = 1,    /* S explicit: BASELINE.json throughput shape D=4                        */
  MAGI_MODEL_SIRW = 2,     /* test_magi_script.py:19-45 (S,I,R,W), theta=(beta,phi,xi,chi,kappa)    */
  MAGI_MODEL_LORENZ96 = 3, /* D=10, theta=(F)                                                       */
  MAGI_MODEL_USER = 100    /* a user-supplied f_vec (magi_v2.py:32-33), traced and compiled at run time into
                              its own library built from csrc/posterior_wide.cu (magi_v2_b200/tracing.py)    */
} magi_model_t;

typedef void* magi_stream_t; /* cudaStream_t */

/* ---- introspection ------------------------------------------------------------------------- */
MAGI_API int magi_b200_abi_version(void);
/* D and P of a compiled-in model; returns 0 or -1 (unknown model). Host-only. */
MAGI_API int magi_b200_model_dims(int model_id, int* D, int* P);
/* Human-readable text for a status code (static storage). Host-only. */
MAGI_API const char* magi_b200_status_string(int status);

/* ---- (1) Matern covariance blocks --------------------------------------------------------------
 * Replaces magi_v2.py:781-815 (`_build_matrices`, the kvp/elementwise part), for all datasets and
 * components at once:  C = kappa, Cp = d kappa/d s ("p_Kappa", antisymmetric, diag 0),
 * Cpp = d^2 kappa / d s d t ("Kappa_pp"), Matern smoothness nu (2.01 in the reference).
 *   I    [B, n] grid points (I_batch_stride = n) or one shared grid [n] (I_batch_stride = 0)
 *   phi1 [B*D], phi2 [B*D]
 *   flags  MAGI_COV_UNIFORM_GRID: the caller asserts the grid is evenly spaced; lags are then taken
 *          as |i-j| * (I[n-1]-I[0])/(n-1) and only n Bessel evaluations per matrix are made
 *          (Toeplitz); without it l = |I_i - I_j| exactly as magi_v2.py:784
 *   C, Cp, Cpp  [B*D, n, n]  (outputs; any may be NULL to skip)                                   */
#define MAGI_COV_UNIFORM_GRID 1
MAGI_API int magi_b200_cov_build(const double* I, int64_t I_batch_stride, const double* phi1, const double* phi2,
                                 double nu, int B, int D, int n, int flags, double* C, double* Cp, double* Cpp,
                                 magi_stream_t stream);

/* ---- (2) factorise + derive ---------------------------------------------------------------------
 * Replaces magi_v2.py:818-820 (pinv(Kappa), m = p_Kappa Kappa^-1, K = Kappa_pp - p_Kappa Kappa^-1 Kappa_p),
 * :126-128 (C^-1, K^-1) and :271-274 (band_part), with Cholesky factorisations instead of SVD
 * pseudo-inverses.  nmat = B*D matrices of order n.
 *   band   < 0: no banding;  >= 0: zero |i-j| > band on Cinv, m, Kinv (as tf.linalg.band_part)
 *   jitter : added to diag(C) and diag(K) before factorising (0 in the reference)
 *   K      optional output (may be NULL)
 *   info   [nmat] int32: 0 ok; k>0: C not positive definite at pivot k; -k: K not PD at pivot k
 *   workspace: magi_b200_factor_workspace_bytes(nmat, n) bytes                                    */
MAGI_API size_t magi_b200_factor_workspace_bytes(int nmat, int n);
MAGI_API int magi_b200_factor_derive(const double* C, const double* Cp, const double* Cpp, int nmat, int n,
                            int band, double jitter, double* Cinv, double* m, double* Kinv, double* K,
                            int32_t* info, void* workspace, size_t workspace_bytes, magi_stream_t stream);

/* Inverse and log-determinant of nmat symmetric positive definite matrices of order n (the Gaussian-process
 * covariance S = phi1 R(phi2) + (sigma^2 + jitter) I of the hyper-parameter fit, magi_v2.py:594-597, where TFP
 * factorises inside MultivariateNormalTriL / GaussianProcess.log_prob) -- the same blocked Cholesky, triangular
 * inverse and L^-T L^-1 product as `magi_b200_factor_derive`.
 *   A [nmat,n,n] (lower triangle read), Ainv [nmat,n,n] (full symmetric), logdet [nmat], info [nmat] (0 ok; k > 0:
 *   not positive definite at pivot k); workspace: magi_b200_factor_workspace_bytes(nmat, n) bytes                   */
MAGI_API int magi_b200_spd_inverse(const double* A, int nmat, int n, double* Ainv, double* logdet, int32_t* info,
                                   void* workspace, size_t workspace_bytes, magi_stream_t stream);

/* ---- (3a) capture the constants of the log-posterior ---------------------------------------------
 * Replaces the closure capture at magi_v2.py:294-296: re-lays C^-1, m, K^-1 [B, D, n, n] into the
 * sampler's device format (opaque; per (b,d): sym(C^-1) | m | sym(K^-1), each padded with zeros to
 * np = 8*ceil(n/8) and stored as 8x8 tiles (element order inside a tile: the bank-conflict-free swizzle of
 * csrc/common.cuh, ABI version 3) so that a warp streams 8 matrix rows as one contiguous
 * run; sym(A) = (A + A^T)/2 so that value AND gradient of x^T A x are those of the possibly
 * non-symmetric A the reference holds).  Done once per fit; `packed` needs
 * magi_b200_packed_bytes(B, D, n) bytes.                                                           */
MAGI_API size_t magi_b200_packed_bytes(int B, int D, int n);
MAGI_API int magi_b200_pack_matrices(const double* Cinv, const double* m, const double* Kinv, int B, int D, int n,
                            void* packed, magi_stream_t stream);

/* Problem constants shared by (3b)-(3d).  Everything the reference's `unnormalized_log_prob`
 * closes over (magi_v2.py:294-300). */
typedef struct {
  int model_id;          /* magi_model_t */
  int B, R, n, D, P;     /* datasets, chains per dataset, grid size, components, parameters */
  const void* packed;    /* from magi_b200_pack_matrices */
  const double* mu;      /* [B, D]   mu_ds (:114) */
  const double* y;       /* [B, n, D] observations on the grid, any value where unobserved (:100) */
  const uint8_t* mask;   /* [B, n, D] 1 = observed (replaces not_nan_idxs, :96) */
  const double* N_ds;    /* [B, D]   non-NaN raw observation counts (:53) */
  const double* beta;    /* [B]      D*n / sum(N_ds) (:89) */
  const double* LB;      /* [B, D]   sigma_sqs_LB (:299-300) */
  int band;              /* bandsize the packed matrices were banded with (magi_v2.py:271-274): entries with
                            |i-j| > band are exactly zero and their tiles are not read; < 0 = dense */
} magi_problem_t;

/* ---- (3b) log-posterior and analytic gradient ------------------------------------------------------
 * Replaces `unnormalized_log_prob` (magi_v2.py:308-348) plus the reverse-mode gradient TFP takes
 * of it, for B*R chains at once.
 *   X [B,R,n,D], sig_pre [B,R,D], th_pre [B,R,P], beta_temp [B,R]
 *   lp [B,R], gX [B,R,n,D], gsig [B,R,D], gth [B,R,P]  (outputs; gradient of lp)
 *   workspace: magi_b200_sampler_workspace_bytes(prob) bytes (may be NULL if that is 0)           */
MAGI_API size_t magi_b200_sampler_workspace_bytes(const magi_problem_t* prob);
MAGI_API int magi_b200_logpost_grad(const magi_problem_t* prob, const double* X, const double* sig_pre,
                           const double* th_pre, const double* beta_temp, double* lp, double* gX,
                           double* gsig, double* gth, void* workspace, size_t workspace_bytes,
                           magi_stream_t stream);

/* ---- (3c) leapfrog trajectory with caller-supplied momenta -----------------------------------------
 * Replaces TFP's SimpleLeapfrogIntegrator as driven by NoUTurnSampler (magi_v2.py:360-364):
 * n_steps of  p += eps/2 grad; z += eps p; p += eps/2 grad  on all three state parts, identity mass,
 * one scalar step size per chain.  State and momenta are updated in place.
 *   eps [B,R], beta_temp [B,R];  lp_out [B,R] log-posterior at the end point (may be NULL)        */
MAGI_API int magi_b200_leapfrog(const magi_problem_t* prob, double* X, double* sig_pre, double* th_pre,
                       double* pX, double* psig, double* pth, const double* eps,
                       const double* beta_temp, int n_steps, double* lp_out, void* workspace,
                       size_t workspace_bytes, magi_stream_t stream);

/* ---- (3d) HMC sampler ------------------------------------------------------------------------------
 * Replaces `tfp.mcmc.sample_chain` over LogAnnealedNUTS(DualAveragingStepSizeAdaptation(NUTS))
 * (magi_v2.py:357-396, :833-889) with fixed-length HMC transitions: per iteration `it` (global
 * index iter0 + i) beta_temp = max(1/log(it + 2), min_temp) (or fixed_beta_temp if > 0), momenta
 * from Philox4x32-10 keyed by (seed; pair, chain_id0 + b*R + r, it), n_leapfrog steps, Metropolis
 * accept, dual averaging of the per-chain step size while it < num_adapt.
 *   state X, sig_pre, th_pre: in/out.   eps [B,R] in/out.
 *   da_state [B,R,4] in/out: (error_sum, log_averaging_step, log_shrinkage_target, step count)
 *   outputs (each may be NULL): th_samps [n_iter,B,R,P] = softplus(th_pre), sig_samps [n_iter,B,R,D]
 *   = softplus(sig_pre)+LB (:418-419), X_samps [n_iter,B,R,n,D], X_sum / X_sumsq [B,R,n,D] running
 *   sums over iterations >= accum_from, accept_prob [n_iter,B,R], lp_trace [n_iter,B,R]           */
typedef struct {
  int n_iter, n_leapfrog;
  int iter0;             /* global index of the first iteration (temperature schedule, RNG counter) */
  int num_adapt;         /* dual averaging active while global iteration < num_adapt */
  int accum_from;        /* accumulate X_sum/X_sumsq for global iterations >= accum_from */
  double min_temp;       /* 0.1 in the reference (:357, :841) */
  double fixed_beta_temp; /* > 0: use this temperature instead of the schedule */
  double target_accept;  /* 0.75 (:366) */
  uint64_t seed;
  uint32_t chain_id0;    /* global id of chain (b=0, r=0): makes draws independent of sharding */
} magi_hmc_config_t;

MAGI_API int magi_b200_hmc_run(const magi_problem_t* prob, const magi_hmc_config_t* cfg, double* X,
                      double* sig_pre, double* th_pre, double* eps, double* da_state, double* th_samps,
                      double* sig_samps, double* X_samps, double* X_sum, double* X_sumsq,
                      double* accept_prob, double* lp_trace, void* workspace, size_t workspace_bytes,
                      magi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGI_B200_H */
