"""Host-side mirror of the reference's public class ``MAGI_v2`` (magi_v2.py:20-425): same constructor,
``initial_fit`` / ``predict`` / ``update_kernel_matrices`` signatures, same attribute names and the same
result dictionary, with the posterior-evaluation hot path (covariance build, factorisation,
log-posterior + gradient, leapfrog/HMC) running as sm_100a CUDA kernels through libmagi_b200.so.

Differences from the reference, all stated in DESIGN.md:
  * ``f_vec`` must resolve to an ODE system compiled into the library (name, OdeModel, or a callable that
    matches one when probed with numpy inputs) -- see models.py;
  * ``predict`` runs the reference's sampler stack by default (``sampler="nuts"``: NUTS trees, dual averaging, log
    annealing; nuts.py); ``sampler="hmc"`` is fixed-length HMC with the same leapfrog, tempering schedule and
    adaptation, run entirely inside one fused kernel; ``predict`` can run many chains at once (``n_chains``) and is
    seedable (the reference is unseeded);
  * ``initial_fit`` reproduces the reference's theta initialisation AS WRITTEN (magi_v2.py:155-156 reshapes the [n, D]
    right-hand side to (D, n, 1) instead of transposing it); ``MAGI_v2.THETA_INIT_LAYOUT = "transpose"`` selects the
    layout the objective's comment intends;
  * C^-1 and K^-1 come from Cholesky factorisations instead of SVD pseudo-inverses;
A CUDA device is required: there is no CPU fallback."""
from __future__ import annotations

import time
from typing import Callable, Optional, Union

import numpy as np

from . import models as _models


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("magi_v2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


class MAGI_v2:
    """Inputs (magi_v2.py:22-33): D_thetas, ts_obs [N], X_obs [N,D] (NaN = missing), bandsize (None or
    positive int), f_vec (see module docstring)."""

    NU = 2.01                      # magi_v2.py:125
    ADAM_LR = 0.01                 # :161
    THETA_INIT_ITERS = 10000       # :161
    THETA_INIT_LAYOUT = "reference"   # :155-156 as written (a reshape); "transpose" = the layout of :335

    def __init__(self, D_thetas: int, ts_obs: np.ndarray, X_obs: np.ndarray, bandsize: Union[int, None],
                 f_vec: Union[str, Callable, _models.OdeModel], device: str = "cuda:0", jit: Union[bool, None] = None):
        self.D_thetas = D_thetas
        self.BANDSIZE = bandsize
        self.ts_obs = np.asarray(ts_obs, dtype=np.float64)
        self.X_obs = np.asarray(X_obs, dtype=np.float64)
        self.N, self.D = self.X_obs.shape
        # observed vs completely unobserved components (:45-50)
        self.observed_indicators = (~np.isnan(self.X_obs)).mean(axis=0) > 0
        self.observed_components = np.arange(self.D)[self.observed_indicators]
        self.D_observed = len(self.observed_components)
        self.unobserved_components = np.setdiff1d(np.arange(self.D), self.observed_components)
        self.D_unobserved = len(self.unobserved_components)
        self.proper_order = np.argsort(np.concatenate([self.observed_components, self.unobserved_components]))
        self.N_ds = (~np.isnan(self.X_obs)).sum(axis=0)                                   # :53
        self.I, self.X_obs_discret = None, None
        self.beta, self.mag_I = None, None
        self.not_nan_idxs, self.not_nan_cols = None, None
        self.y_tau_ds_observed = None
        self.X_interp_obs, self.X_interp_unobs = None, None
        self.phi1s = np.full((self.D,), np.nan)
        self.phi2s = np.full((self.D,), np.nan)
        self.sigma_sqs_init = np.full((self.D,), np.nan)
        self.Xhat_init, self.thetas_init = None, None
        self.mu_ds = np.full((self.D,), np.nan)
        self.C_d_invs, self.m_ds, self.K_d_invs = None, None, None
        self.f_vec = f_vec
        # jit: None = a callable that reproduces a compiled-in system uses it, any other callable is traced and compiled
        # at run time (tracing.py); True = always trace and compile; False = compiled-in systems only
        self.model = _models.resolve(f_vec, self.D, D_thetas, jit=jit)
        if (self.model.D, self.model.P) != (self.D, D_thetas):
            raise ValueError(f"model {self.model.name} has D={self.model.D}, P={self.model.P}; "
                             f"data has D={self.D}, D_thetas={D_thetas}")
        self.device = device
        self.factor_info = None

    # ------------------------------------------------------------------------------------------
    def initial_fit(self, discretization: int, verbose=False, hparams: Optional[dict] = None, seed=None):
        """magi_v2.py:82-277.  ``hparams`` = {"phi1s", "phi2s", "sigma_sqs"} (observed components) skips their GP
        hyper-parameter fit (the reference lets the user overwrite the fitted values, :76-80).  ``seed`` seeds the
        random start of completely unobserved components (the reference draws it unseeded, :224-227)."""
        self.I, self.X_obs_discret = self._discretize(self.ts_obs, self.X_obs, discretization)      # :85
        self.mag_I = self.I.shape[0]
        self.beta = float((self.D * self.mag_I) / self.N_ds.sum())                                   # :89
        self.not_nan_idxs = np.where(~np.isnan(self.X_obs_discret).flatten())[0]                     # :96
        self.not_nan_cols = self.not_nan_idxs % self.D                                               # :97
        self.y_tau_ds_observed = self.X_obs_discret.reshape(-1)[self.not_nan_idxs]                   # :100
        self.X_interp_obs = self._linear_interpolate(self.X_obs_discret[:, self.observed_indicators])  # :105
        if hparams is None:
            hparams = self._fit_kernel_hparams(I=self.I, X_filled=self.X_interp_obs, verbose=verbose)  # :106
        self.phi1s[self.observed_indicators] = np.asarray(hparams["phi1s"], dtype=np.float64)
        self.phi2s[self.observed_indicators] = np.asarray(hparams["phi2s"], dtype=np.float64)
        self.sigma_sqs_init[self.observed_indicators] = np.asarray(hparams["sigma_sqs"], dtype=np.float64)
        self.Xhat_init = self.X_obs_discret.copy()
        self.Xhat_init[:, self.observed_indicators] = self.X_interp_obs
        self.mu_ds[self.observed_indicators] = self.X_interp_obs.mean(axis=0)                        # :114
        if self.D_unobserved == 0:
            # kernel matrices on the device: replaces the per-component loop :122-128.  The reference fits thetas_init
            # on the UN-banded matrices (band_part only runs at :271-274), so the band is applied afterwards.
            self._device_kernel_matrices(band=None)
            # theta initialisation (:132-179): Adam on the t2-only objective from theta = 1
            self.thetas_init = self._fit_thetas_init()
            self._apply_band()                                                                       # :271-274
        else:
            # :182-268 -- (thetas_init, X_unobs) jointly by gradient matching, then the GP hyper-parameters of the
            # unobserved components from the fitted trajectories, then every kernel matrix in one device call
            from .init_fit import fit_unobserved
            X_smoothed_obs = self.cv_cubic_smoother(self.I, self.X_interp_obs)                       # :192
            self.X_interp_unobs, self.thetas_init, l0, l1 = fit_unobserved(
                self.model, self.I, X_smoothed_obs, self.observed_components, self.unobserved_components,
                self.X_interp_obs, num_iters=self.THETA_INIT_ITERS, lr=self.ADAM_LR, seed=seed, device=self.device)
            if verbose:
                print(f"Fitting X_unobs and theta: gradient-matching loss {l0:.4g} -> {l1:.4g}")
            hp_u = self._fit_kernel_hparams(I=self.I, X_filled=self.X_interp_unobs, verbose=verbose)  # :253
            self.phi1s[self.unobserved_components] = hp_u["phi1s"]
            self.phi2s[self.unobserved_components] = hp_u["phi2s"]
            self.sigma_sqs_init[self.unobserved_components] = hp_u["sigma_sqs"]
            self.Xhat_init[:, self.unobserved_components] = self.X_interp_unobs
            self.mu_ds[self.unobserved_components] = self.X_interp_unobs.mean(axis=0)                # :260
            self._device_kernel_matrices()                                                           # :262-274
        self.Xhat_init = self.cv_cubic_smoother(self.I, self.Xhat_init)                              # :277

    def _device_kernel_matrices(self, band="default"):
        torch = _require_cuda()
        from . import ops
        dev = torch.device(self.device)
        I = torch.as_tensor(self.I.ravel(), dtype=torch.float64, device=dev)
        p1 = torch.as_tensor(self.phi1s[None], dtype=torch.float64, device=dev)
        p2 = torch.as_tensor(self.phi2s[None], dtype=torch.float64, device=dev)
        C, Cp, Cpp = ops.cov_build(I, p1, p2, self.NU, False)
        if band == "default":
            band = self.BANDSIZE
        Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, -1 if band is None else int(band), 0.0)
        self.factor_info = info[0].cpu().numpy()
        if np.any(self.factor_info != 0):
            raise np.linalg.LinAlgError(f"covariance not positive definite, info={self.factor_info}")
        self._dev_mats = (Cinv, m, Kinv)
        self.C_d_invs, self.m_ds, self.K_d_invs = (a[0].cpu().numpy() for a in (Cinv, m, Kinv))

    def _apply_band(self):
        """tf.linalg.band_part(., BANDSIZE, BANDSIZE) on C^-1, K^-1, m (magi_v2.py:271-274)."""
        if self.BANDSIZE is None:
            return
        import torch
        n = self.mag_I
        i = torch.arange(n, device=self._dev_mats[0].device)
        keep = (i[:, None] - i[None, :]).abs() <= int(self.BANDSIZE)
        self._dev_mats = tuple(torch.where(keep, a, torch.zeros_like(a)) for a in self._dev_mats)
        self.C_d_invs, self.m_ds, self.K_d_invs = (a[0].cpu().numpy() for a in self._dev_mats)

    def _fit_thetas_init(self):
        """magi_v2.py:132-179: minimise t2(theta) = sum_d r_d^T K_d^-1 r_d with Adam (lr 0.01, 10 000 steps from
        theta = 1) at the linearly interpolated Xhat_init, on the un-banded matrices.  Every compiled-in right-hand side
        is affine in theta, so t2 is the quadratic theta^T A theta - 2 b^T theta + c; Adam's iterates are reproduced
        on that quadratic.

        THETA_INIT_LAYOUT = "reference" (default) reproduces :155-156 as written: the [n, D] output of f_vec is
        ``tf.reshape``d -- not transposed -- to (D, n, 1), which interleaves components and grid points (on the
        vignette it drives every theta negative, so sampling starts from softplus(-5) = 0.0067, :381-382).
        "transpose" uses the layout unnormalized_log_prob uses (:335), i.e. what the objective's comment intends."""
        X, I = self.Xhat_init, self.I
        P, D, n = self.D_thetas, self.D, self.mag_I
        xc = (X - self.mu_ds).T                                            # [D,n]
        mx = np.einsum("dij,dj->di", self.m_ds, xc)
        if not self.model.affine_in_theta:
            return self._fit_thetas_init_general(mx)
        f0 = self.model.f_vec(I, X, np.zeros(P))                           # [n,D] part independent of theta
        F = self.model.dtheta(I, X, np.zeros(P))                           # [n,D,P]
        if self.THETA_INIT_LAYOUT == "reference":
            f0, F = f0.reshape(D, n), F.reshape(D, n, P)                   # :155-156
        elif self.THETA_INIT_LAYOUT == "transpose":
            f0, F = f0.T, np.transpose(F, (1, 0, 2))
        else:
            raise ValueError("THETA_INIT_LAYOUT must be 'reference' or 'transpose'")
        r0 = f0 - mx
        KF = np.einsum("dij,djk->dik", self.K_d_invs, F)
        KTF = np.einsum("dji,djk->dik", self.K_d_invs, F)
        A = np.einsum("dik,dil->kl", F, KF)
        b2 = np.einsum("di,dik->k", r0, KF) + np.einsum("di,dik->k", r0, KTF)
        th = np.ones(P)
        m1, v1 = np.zeros(P), np.zeros(P)
        b1, b2a, eps = 0.9, 0.999, 1e-7                                    # tf_keras Adam defaults
        for t in range(1, self.THETA_INIT_ITERS + 1):
            g = (A + A.T) @ th + b2
            m1 = b1 * m1 + (1 - b1) * g
            v1 = b2a * v1 + (1 - b2a) * g * g
            lr_t = self.ADAM_LR * np.sqrt(1 - b2a ** t) / (1 - b1 ** t)
            th = th - lr_t * m1 / (np.sqrt(v1) + eps)
        return th

    def _fit_thetas_init_general(self, mx):
        """The same Adam run for a right-hand side that is not affine in theta (a traced user system): gradient of t2
        by the chain rule through the traced d f / d theta."""
        X, I = self.Xhat_init, self.I
        P, D, n = self.D_thetas, self.D, self.mag_I
        lay = (lambda a: a.reshape(D, n, *a.shape[2:])) if self.THETA_INIT_LAYOUT == "reference" else \
            (lambda a: np.moveaxis(a, 1, 0))
        SK = self.K_d_invs + np.transpose(self.K_d_invs, (0, 2, 1))
        th = np.ones(P)
        m1, v1 = np.zeros(P), np.zeros(P)
        b1, b2a, eps = 0.9, 0.999, 1e-7
        for t in range(1, self.THETA_INIT_ITERS + 1):
            r = lay(self.model.f_vec(I, X, th)) - mx                       # [D,n]
            F = lay(self.model.dtheta(I, X, th))                           # [D,n,P]
            g = np.einsum("dik,di->k", F, np.einsum("dij,dj->di", SK, r))
            m1 = b1 * m1 + (1 - b1) * g
            v1 = b2a * v1 + (1 - b2a) * g * g
            th = th - (self.ADAM_LR * np.sqrt(1 - b2a ** t) / (1 - b1 ** t)) * m1 / (np.sqrt(v1) + eps)
        return th

    # ------------------------------------------------------------------------------------------
    def predict(self, num_results: int = 1000, num_burnin_steps: int = 1000, sigma_sqs_LB=None, verbose=False,
                n_chains: int = 1, n_leapfrog: int = 32, seed: int = 0, step_size: float = None,
                init_jitter: float = 0.0, sampler: str = "nuts", max_tree_depth: int = 10,
                beta_temp: Optional[float] = None, cached_target: bool = True):
        """magi_v2.py:286-425.  Returns the reference's result dictionary; with n_chains > 1 the sample
        arrays gain a leading chain axis.  sampler = "hmc": fixed-length trajectories, the whole chain inside the
        fused CUDA kernel; sampler = "nuts": the reference's sampler (No-U-Turn trees, `nuts.py`) with one launch
        of the log-posterior + gradient kernel per leapfrog step.  beta_temp = None follows the reference's schedule
        max(1 / log(step + 2), 0.1) (:833-835, which never returns to 1); a number fixes the temperature (1.0 = the
        untempered posterior).  cached_target (NUTS only): start every transition from the target value / gradient of
        the previous step's temperature, as TFP's kernel results do under the reference's wrapper (nuts.nuts_run_)."""
        if sampler not in ("hmc", "nuts"):
            raise ValueError("sampler must be 'hmc' or 'nuts'")
        torch = _require_cuda()
        from . import ops
        assert ~np.any(np.isnan(self.Xhat_init)), "Please make sure Xhat_init does not have NaNs."
        assert ~np.any(np.isnan(self.sigma_sqs_init)), "Please make sure sigma_sqs_init does not have NaNs."
        assert ~np.any(np.isnan(self.thetas_init)), "Please make sure thetas_init does not have NaNs."
        if sigma_sqs_LB is None:
            sigma_sqs_LB = ((self.Xhat_init.std(axis=0)) * 0.01) ** 2                                # :299-300
        sigma_sqs_LB = np.asarray(sigma_sqs_LB, dtype=np.float64)
        # initial state (:373-383)
        sigma_sqs_pre_init = np.full_like(self.sigma_sqs_init, -5.0)
        ok = self.sigma_sqs_init > sigma_sqs_LB
        sigma_sqs_pre_init[ok] = np.log(np.exp((self.sigma_sqs_init - sigma_sqs_LB)[ok]) - 1.0)
        thetas_pre_init = np.full_like(self.thetas_init, -5.0)
        ok = self.thetas_init > 0.0
        thetas_pre_init[ok] = np.log(np.exp(self.thetas_init[ok]) - 1.0)

        dev = torch.device(self.device)
        T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        n, D, P, R = self.mag_I, self.D, self.D_thetas, int(n_chains)
        Cinv, m, Kinv = (T(a[None]) for a in (self.C_d_invs, self.m_ds, self.K_d_invs))
        packed = ops.pack_matrices(Cinv, m, Kinv)
        y = np.zeros(n * D); mask = np.zeros(n * D, dtype=np.uint8)
        y[self.not_nan_idxs] = self.y_tau_ds_observed
        mask[self.not_nan_idxs] = 1
        prob = ops.PosteriorProblem(self.model if self.model.lib_path else self.model.name, packed,
                                    mu=T(self.mu_ds[None]), y=T(y.reshape(1, n, D)),
                                    mask=T(mask.reshape(1, n, D), torch.uint8),
                                    N_ds=T(self.N_ds[None].astype(np.float64)), beta=T(np.array([self.beta])),
                                    LB=T(sigma_sqs_LB[None]), n=n, band=self.BANDSIZE)
        rng = np.random.default_rng(seed)
        X0 = np.repeat(self.Xhat_init[None, None], R, axis=1)
        if init_jitter > 0.0:
            X0 = X0 + init_jitter * self.Xhat_init.std(axis=0) * rng.standard_normal(X0.shape)
        X = T(X0)
        s = T(np.repeat(sigma_sqs_pre_init[None, None], R, axis=1))
        tau = T(np.repeat(thetas_pre_init[None, None], R, axis=1))
        eps0 = 0.1 if step_size is None else float(step_size)                                       # :364
        eps = torch.full((1, R), eps0, dtype=torch.float64, device=dev)
        da = torch.zeros((1, R, 4), dtype=torch.float64, device=dev)
        da[..., 2] = float(np.log(10.0 * eps0))
        num_adapt = int(0.8 * num_burnin_steps)                                                     # :365
        if sampler == "nuts":
            return self._predict_nuts(prob, X, s, tau, eps, da, num_results, num_burnin_steps, num_adapt, seed,
                                      sigma_sqs_LB, max_tree_depth, verbose, beta_temp, cached_target)
        if verbose:
            print("Starting HMC posterior sampling ...")
        start = time.time()
        fbt = 0.0 if beta_temp is None else float(beta_temp)
        burn = prob.hmc_run_(X, s, tau, eps, da, n_iter=num_burnin_steps, n_leapfrog=n_leapfrog, iter0=0,
                             num_adapt=num_adapt, seed=seed, keep_theta=False, keep_sigma=False, fixed_beta_temp=fbt)
        out = prob.hmc_run_(X, s, tau, eps, da, n_iter=num_results, n_leapfrog=n_leapfrog, iter0=num_burnin_steps,
                            num_adapt=num_adapt, seed=seed, keep_X=True, fixed_beta_temp=fbt)
        torch.cuda.synchronize(dev)
        end = time.time()
        minutes = np.round((end - start) / 60, 2)
        if verbose:
            print(f"Finished sampling in {minutes} minutes.")
        sq = (lambda a: a[:, 0, 0]) if R == 1 else (lambda a: np.moveaxis(a[:, 0], 1, 0))
        X_s = sq(out["X_samps"].cpu().numpy())
        sig_s = sq(out["sigma_sqs_samps"].cpu().numpy())
        th_s = sq(out["thetas_samps"].cpu().numpy())
        kernel_results = {"accept_prob": sq(out["accept_prob"].cpu().numpy()),
                          "target_log_prob": sq(out["lp"].cpu().numpy()),
                          "burnin_accept_prob": sq(burn["accept_prob"].cpu().numpy()),
                          "step_size": eps[0].cpu().numpy(), "n_leapfrog": n_leapfrog, "sampler": "hmc"}
        return {"phi1s": self.phi1s, "phi2s": self.phi2s, "Xhat_init": self.Xhat_init,
                "sigma_sqs_init": self.sigma_sqs_init, "thetas_init": self.thetas_init, "I": self.I,
                "X_samps": X_s, "sigma_sqs_samps": sig_s, "thetas_samps": th_s,
                "kernel_results": kernel_results,
                "sample_results": [X_s, np.log(np.expm1(np.maximum(sig_s - sigma_sqs_LB, 1e-300))),
                                   np.log(np.expm1(th_s))],
                "minutes_elapsed": minutes}

    def _predict_nuts(self, prob, X, s, tau, eps, da, num_results, num_burnin_steps, num_adapt, seed, sigma_sqs_LB,
                      max_tree_depth, verbose, beta_temp=None, cached_target=True):
        """The reference's sampler stack (magi_v2.py:357-396): NUTS in dual averaging in the annealing wrapper."""
        import torch
        from . import nuts
        R, n, D, P = X.shape[1], self.mag_I, self.D, self.D_thetas
        z = nuts.pack_state(X, s, tau)
        e, d = eps.reshape(-1).clone(), da.reshape(-1, 4).clone()
        eng = nuts.FusedLeafEngine(prob, R)
        if verbose:
            print("Starting NUTS posterior sampling ...")
        start = time.time()
        burn = nuts.nuts_run_(z, e, d, None, n_iter=num_burnin_steps, iter0=0, num_adapt=num_adapt, seed=seed,
                              max_tree_depth=max_tree_depth, leaf_engine=eng, fixed_beta_temp=beta_temp,
                              cached_target=cached_target)
        keep = torch.empty((num_results,) + tuple(z.shape), dtype=torch.float64, device=z.device)
        out = nuts.nuts_run_(z, e, d, None, n_iter=num_results, iter0=num_burnin_steps, num_adapt=num_adapt, seed=seed,
                             max_tree_depth=max_tree_depth, leaf_engine=eng, fixed_beta_temp=beta_temp,
                             cached_target=cached_target, on_sample=lambda it, zz, info: keep[it].copy_(zz))
        torch.cuda.synchronize(z.device)
        minutes = np.round((time.time() - start) / 60, 2)
        if verbose:
            print(f"Finished sampling in {minutes} minutes.")
        k = keep.cpu().numpy()                                               # [num_results, R, S]
        sq = (lambda a: a[:, 0]) if R == 1 else (lambda a: np.moveaxis(a, 1, 0))
        X_pre, s_pre, t_pre = sq(k[..., :n * D].reshape(num_results, R, n, D)), sq(k[..., n * D:n * D + D]), \
            sq(k[..., n * D + D:])
        kernel_results = {"accept_prob": sq(out["accept_prob"].cpu().numpy()),
                          "target_log_prob": sq(out["lp"].cpu().numpy()),
                          "leapfrogs_taken": sq(out["n_leapfrog"].cpu().numpy()),
                          "has_divergence": sq(out["diverged"].cpu().numpy()),
                          "burnin_accept_prob": sq(burn["accept_prob"].cpu().numpy()),
                          "burnin_leapfrogs_taken": sq(burn["n_leapfrog"].cpu().numpy()),
                          "step_size": e.cpu().numpy(), "sampler": "nuts"}
        return {"phi1s": self.phi1s, "phi2s": self.phi2s, "Xhat_init": self.Xhat_init,
                "sigma_sqs_init": self.sigma_sqs_init, "thetas_init": self.thetas_init, "I": self.I,
                "X_samps": X_pre, "sigma_sqs_samps": np.log(np.exp(s_pre) + 1.0) + sigma_sqs_LB,        # :418
                "thetas_samps": np.log(np.exp(t_pre) + 1.0),                                             # :419
                "kernel_results": kernel_results, "sample_results": [X_pre, s_pre, t_pre],
                "minutes_elapsed": minutes}

    # ------------------------------------------------------------------------------------------
    def update_kernel_matrices(self, I_new, phi1s_new, phi2s_new):
        """magi_v2.py:433-462 (forecasting: new grid / hyper-parameters, same observations)."""
        self.I = np.asarray(I_new, dtype=np.float64).reshape(-1, 1)
        self.phi1s, self.phi2s = np.array(phi1s_new, dtype=np.float64), np.array(phi2s_new, dtype=np.float64)
        self.mag_I = self.I.shape[0]
        self.beta = float((self.D * self.mag_I) / self.N_ds.sum())
        self._device_kernel_matrices()

    # ------------------------------------------------------------------------------------------
    # helper functions (host side, numpy/scipy as in the reference)
    # ------------------------------------------------------------------------------------------
    def _discretize(self, ts_obs, X_obs, discretization):
        """magi_v2.py:475-498."""
        ts_obs = ts_obs.flatten()
        assert ts_obs.shape[0] == X_obs.shape[0], \
            "Please make sure there are equal numbers of observations in ts_obs and X_obs."
        N, D = X_obs.shape
        stride = 2 ** discretization
        N_discret = stride * (N - 1) + 1
        I = np.full((N_discret,), np.nan)
        X_obs_discret = np.full((N_discret, D), np.nan)
        I[::stride] = ts_obs
        indices = np.arange(len(I))
        have = ~np.isnan(I)
        I = np.interp(x=indices, xp=indices[have], fp=I[have]).reshape(-1, 1)
        X_obs_discret[::stride] = X_obs
        return I, X_obs_discret

    def _linear_interpolate(self, X_partial):
        """magi_v2.py:509-527."""
        X_interp = X_partial.copy()
        indices = np.arange(X_partial.shape[0])
        for d in range(X_partial.shape[1]):
            col = X_partial[:, d]
            if np.any(np.isnan(col)):
                have = ~np.isnan(col)
                X_interp[:, d] = np.interp(x=indices, xp=indices[have], fp=col[have])
        return X_interp

    def _fit_kernel_hparams(self, I, X_filled, verbose=False):
        """magi_v2.py:538-691 -- GP hyper-parameter fit (Fourier-informed prior + 1000 Adam steps on the
        GP marginal likelihood), batched on the device."""
        from .hparams import fit_kernel_hparams
        out = fit_kernel_hparams(np.asarray(I).ravel(), X_filled[None], device=self.device, verbose=verbose)
        return {k: v[0] for k, v in out.items()}

    def _build_matrices(self, I, phi1, phi2, v=2.01):
        """magi_v2.py:774-823 on the device: returns (C_d, m_d, K_d) as numpy arrays."""
        torch = _require_cuda()
        from . import ops
        dev = torch.device(self.device)
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        C, Cp, Cpp = ops.cov_build(T(np.asarray(I).ravel()), T([[phi1]]), T([[phi2]]), float(v), False)
        _, m, _, K, info = ops.factor_derive(C, Cp, Cpp, -1, 0.0)
        if int(info.abs().max()) != 0:
            raise np.linalg.LinAlgError("covariance not positive definite")
        return C[0, 0].cpu().numpy(), m[0, 0].cpu().numpy(), K[0, 0].cpu().numpy()

    def cv_cubic_smoother(self, I, X_filled):
        """magi_v2.py:695-703."""
        I = I.flatten()
        if I.shape[0] < 10:
            return X_filled
        return np.stack([self.single_cv_cubic_smoother(I, X_filled[:, i]) for i in range(X_filled.shape[1])], axis=1)

    def single_cv_cubic_smoother(self, I, x):
        """magi_v2.py:707-770.  As in the reference, the 5-fold CV over knot counts is evaluated but the
        final spline is fitted with the LAST knot count tried (the loop variable), not the CV optimum
        (:747 vs :750-767); reproduced so that Xhat_init matches."""
        from scipy.interpolate import splev, splrep
        I = I.flatten()
        if I.shape[0] < 10:
            return x
        knot_num = I.shape[0] // 10                       # last value of the reference's loop variable
        if knot_num == 0:
            knot_positions = np.array([])
        else:
            knot_positions = np.linspace(start=I[0], stop=I[-1], num=knot_num + 2)[1:-1]
        tck = splrep(I, x, t=knot_positions, s=0)
        return splev(I, tck)


def logarithmic_temperature_schedule(step, min_temp: float = 0.1):
    """magi_v2.py:833-835."""
    return np.maximum(1.0 / np.log(np.asarray(step, dtype=np.float64) + 2.0), min_temp)
