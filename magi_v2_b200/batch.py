"""Many datasets at once: the batched counterpart of `MAGI_v2.initial_fit` + `MAGI_v2.predict`
(magi_v2.py:82-277, :286-425) for B fully observed datasets on a common time base -- BASELINE.json configs 2
and 4.  Every dataset gets its own kernel hyper-parameters, matrices, theta initialisation and chains; all
of them are sampled by ONE launch of the fused HMC kernel per block of iterations.  With `torch.distributed`
initialised, the datasets are sharded over the ranks (contiguous blocks, all chains of a dataset on one
rank) and the theta / sigma^2 samples are all-gathered at the end -- the only collective.

A CUDA device is required: there is no CPU fallback."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import models as _models
from . import parallel, synth


class MagiBatch:
    """ts_obs [N], X_obs [B, N, D] (NaN = missing entry; every component observed at least once),
    bandsize, model name.  Mirrors the attribute names of the reference class with a leading dataset axis."""

    NU = 2.01
    THETA_INIT_LAYOUT = "reference"   # magi_v2.py:155-156 as written; see MAGI_v2._fit_thetas_init

    def __init__(self, ts_obs: np.ndarray, X_obs: np.ndarray, bandsize: Optional[int], model: str,
                 device: Optional[str] = None):
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise RuntimeError("magi_v2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.model = _models.REGISTRY[model]
        X_obs = np.asarray(X_obs, dtype=np.float64)
        if X_obs.ndim != 3 or X_obs.shape[2] != self.model.D:
            raise ValueError(f"X_obs must be [B, N, {self.model.D}] for model {model}")
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.B_total = X_obs.shape[0]
        if self.B_total < self.world:
            raise ValueError(f"{self.B_total} datasets cannot be sharded over {self.world} ranks: every rank needs at "
                             "least one dataset (run with fewer ranks)")
        self.lo, self.hi = parallel.shard_range(self.B_total, self.rank, self.world)
        self.ts_obs = np.asarray(ts_obs, dtype=np.float64)
        self.X_obs = X_obs[self.lo:self.hi]                      # this rank's datasets
        self.BANDSIZE = bandsize
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.prob = None

    @property
    def B(self):
        return self.X_obs.shape[0]

    def initial_fit(self, discretization: int, hparams: Optional[dict] = None, hparam_iters: int = 1000,
                    verbose: bool = False):
        """magi_v2.py:82-277 for every local dataset: grid, interpolation, GP hyper-parameters (fitted on
        the device unless given as {"phi1s", "phi2s", "sigma_sqs"} [B_local, D]), kernel matrices, theta init."""
        from . import hparams as _hp
        c = synth.batch_constants(self.ts_obs, self.X_obs, discretization)
        self.I, self.consts = c["I"], c
        if hparams is None:
            hparams = _hp.fit_kernel_hparams(c["I"], c["Xhat"], device=self.device, num_iters=hparam_iters,
                                             verbose=verbose)
        self.phi1s, self.phi2s = np.asarray(hparams["phi1s"]), np.asarray(hparams["phi2s"])
        self.sigma_sqs_init = np.asarray(hparams["sigma_sqs"])
        self.X_interp = c["Xhat"]                                                         # linear interpolants, :105
        # matrices factorised WITHOUT the band: the reference fits thetas_init on the un-banded m, K^-1 (:132-179) and
        # applies band_part afterwards (:271-274); device_problem bands them before packing
        self.prob, info = synth.device_problem(self.model.name, c["I"], self.phi1s, self.phi2s, c["y"], c["mask"],
                                               c["N_ds"], c["beta"], c["mu"], None, self.BANDSIZE,
                                               self.device, nu=self.NU, uniform_grid=self._uniform_grid(),
                                               keep_matrices=True)
        self.factor_info = info.cpu().numpy()
        if np.any(self.factor_info != 0):
            bad = np.argwhere(self.factor_info != 0)[:5].tolist()
            raise np.linalg.LinAlgError(f"covariance not positive definite for (dataset, component) {bad}")
        self.thetas_init = self._fit_thetas_init()
        self.prob.kept_matrices = None
        self.Xhat_init = self.cv_cubic_smoother(c["I"], self.X_interp)                    # :277
        self.sigma_sqs_LB = (0.01 * self.Xhat_init.std(axis=1)) ** 2                      # :299-300
        self.prob.set_LB(self.sigma_sqs_LB)
        return self

    @staticmethod
    def cv_cubic_smoother(I, X_filled):
        """magi_v2.py:695-770 for all datasets at once.  The reference's cross-validation result is unused: the final
        spline is always the least-squares cubic spline on the LARGEST knot count tried, n // 10 equally spaced
        interior knots (:750-767; see MAGI_v2.single_cv_cubic_smoother), so one batched least-squares fit does it."""
        from scipy.interpolate import make_lsq_spline
        I = np.asarray(I, dtype=np.float64).reshape(-1)
        n = I.shape[0]
        if n < 10:
            return X_filled
        knot_num = n // 10
        interior = np.linspace(I[0], I[-1], knot_num + 2)[1:-1] if knot_num > 0 else np.array([])
        t = np.concatenate([[I[0]] * 4, interior, [I[-1]] * 4])
        B, _, D = X_filled.shape
        y = np.transpose(X_filled, (1, 0, 2)).reshape(n, B * D)
        return np.transpose(make_lsq_spline(I, y, t, k=3)(I).reshape(n, B, D), (1, 0, 2))

    def _uniform_grid(self) -> bool:
        steps = np.diff(np.asarray(self.I, dtype=np.float64))
        return bool(np.allclose(steps, steps[0], rtol=1e-10, atol=0.0))

    def _fit_thetas_init(self, iters: int = 10000, lr: float = 0.01):
        """magi_v2.py:132-179 batched: Adam (lr 0.01, 10 000 steps from theta = 1) on t2(theta), which is
        quadratic in theta for the compiled-in systems; the quadratic's coefficients come from the device
        matrices, the Adam recursion runs on the device for all datasets at once."""
        import torch
        m, Kinv = self.prob.kept_matrices                                    # [B,D,n,n] device, un-banded
        dev, B, P, D = self.device, self.B, self.model.P, self.model.D
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        X = self.X_interp
        n = X.shape[1]
        f0 = np.stack([self.model.f_vec(None, X[b], np.zeros(P)) for b in range(B)])            # [B,n,D]
        F = np.stack([self.model.dtheta(None, X[b], np.zeros(P)) for b in range(B)])            # [B,n,D,P]
        if self.THETA_INIT_LAYOUT == "reference":                                                # :155-156: a reshape
            f0l, Fl = f0.reshape(B, D, n), F.reshape(B, D, n, P)
        elif self.THETA_INIT_LAYOUT == "transpose":                                              # the layout of :335
            f0l, Fl = np.transpose(f0, (0, 2, 1)), np.transpose(F, (0, 2, 1, 3))
        else:
            raise ValueError("THETA_INIT_LAYOUT must be 'reference' or 'transpose'")
        xc = T(np.transpose(X - self.consts["mu"][:, None], (0, 2, 1)))                          # [B,D,n]
        r0 = T(f0l) - torch.einsum("bdij,bdj->bdi", m, xc)
        Ft = T(Fl)                                                                               # [B,D,n,P]
        KF = torch.einsum("bdij,bdjk->bdik", Kinv, Ft)
        KTF = torch.einsum("bdji,bdjk->bdik", Kinv, Ft)
        A = torch.einsum("bdik,bdil->bkl", Ft, KF)
        b2 = torch.einsum("bdi,bdik->bk", r0, KF) + torch.einsum("bdi,bdik->bk", r0, KTF)
        As = A + A.transpose(1, 2)
        th = torch.ones((B, P), dtype=torch.float64, device=dev)
        m1, v1 = torch.zeros_like(th), torch.zeros_like(th)
        b1_, b2_, eps = 0.9, 0.999, 1e-7
        for t in range(1, iters + 1):
            g = torch.einsum("bkl,bl->bk", As, th) + b2
            m1 = b1_ * m1 + (1 - b1_) * g
            v1 = b2_ * v1 + (1 - b2_) * g * g
            th = th - (lr * np.sqrt(1 - b2_ ** t) / (1 - b1_ ** t)) * m1 / (torch.sqrt(v1) + eps)
        return th.cpu().numpy()

    def predict(self, num_results: int = 1000, num_burnin_steps: int = 1000, n_chains: int = 8,
                n_leapfrog: int = 32, seed: int = 0, step_size: float = 0.1, keep_X_mean: bool = True,
                gather: bool = True, sampler: str = "nuts", max_tree_depth: int = 10,
                cached_target: bool = True):
        """magi_v2.py:286-425 for every dataset: returns thetas_samps [B, n_chains, num_results, P],
        sigma_sqs_samps [B, n_chains, num_results, D] (all datasets of all ranks when `gather`), the
        posterior mean / sd of the local trajectories and per-chain acceptance / step sizes.  sampler = "hmc"
        runs whole chains inside the fused kernel; "nuts" is the reference's sampler (`nuts.py`)."""
        import torch
        if self.prob is None:
            raise RuntimeError("call initial_fit() first")
        dev, B, R, D, P = self.device, self.B, int(n_chains), self.model.D, self.model.P
        n = len(self.I)
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        LB = self.sigma_sqs_LB
        s0 = np.where(self.sigma_sqs_init > LB, np.log(np.expm1(np.maximum(self.sigma_sqs_init - LB, 1e-300))), -5.0)
        t0 = np.where(self.thetas_init > 0, np.log(np.expm1(np.maximum(self.thetas_init, 1e-300))), -5.0)  # :373-383
        X = T(np.repeat(self.Xhat_init[:, None], R, axis=1))
        s = T(np.repeat(s0[:, None], R, axis=1))
        tau = T(np.repeat(t0[:, None], R, axis=1))
        eps = torch.full((B, R), float(step_size), dtype=torch.float64, device=dev)
        da = torch.zeros((B, R, 4), dtype=torch.float64, device=dev)
        da[..., 2] = float(np.log(10.0 * step_size))
        num_adapt = int(0.8 * num_burnin_steps)                                                     # :365
        cid0 = self.lo * R                                      # global chain ids: results independent of sharding
        if sampler == "nuts":
            return self._predict_nuts(X, s, tau, eps, da, num_results, num_burnin_steps, num_adapt, seed, cid0,
                                      max_tree_depth, keep_X_mean, gather, cached_target)
        burn = self.prob.hmc_run_(X, s, tau, eps, da, n_iter=num_burnin_steps, n_leapfrog=n_leapfrog, iter0=0,
                                  num_adapt=num_adapt, seed=seed, chain_id0=cid0, keep_theta=False, keep_sigma=False)
        Xsum = torch.zeros((B, R, n, D), dtype=torch.float64, device=dev) if keep_X_mean else None
        Xsq = torch.zeros((B, R, n, D), dtype=torch.float64, device=dev) if keep_X_mean else None
        out = self.prob.hmc_run_(X, s, tau, eps, da, n_iter=num_results, n_leapfrog=n_leapfrog,
                                 iter0=num_burnin_steps, num_adapt=num_adapt, accum_from=num_burnin_steps, seed=seed,
                                 chain_id0=cid0, X_sum=Xsum, X_sumsq=Xsq)
        th, sg = out["thetas_samps"], out["sigma_sqs_samps"]                 # [iter, B, R, .]
        if gather and self.world > 1:
            sizes = parallel.shard_sizes(self.B_total, self.world)
            th = parallel.gather_samples(th, dataset_dim=1, sizes=sizes)
            sg = parallel.gather_samples(sg, dataset_dim=1, sizes=sizes)
        res = {"thetas_samps": th.permute(1, 2, 0, 3).cpu().numpy(), "sigma_sqs_samps": sg.permute(1, 2, 0, 3).cpu().numpy(),
               "accept_prob": out["accept_prob"].mean(dim=0).cpu().numpy(), "step_size": eps.cpu().numpy(),
               "burnin_accept_prob": burn["accept_prob"].mean(dim=0).cpu().numpy(),
               "phi1s": self.phi1s, "phi2s": self.phi2s, "thetas_init": self.thetas_init, "I": self.I,
               "dataset_range": (self.lo, self.hi)}
        if keep_X_mean:
            mean = Xsum.sum(dim=1) / (R * num_results)
            var = Xsq.sum(dim=1) / (R * num_results) - mean ** 2
            res["X_mean"], res["X_sd"] = mean.cpu().numpy(), var.clamp_min(0).sqrt().cpu().numpy()
        return res

    def _predict_nuts(self, X, s, tau, eps, da, num_results, num_burnin_steps, num_adapt, seed, cid0, max_tree_depth,
                      keep_X_mean, gather, cached_target=True):
        import torch
        from . import nuts
        B, R, D, P, n = self.B, X.shape[1], self.model.D, self.model.P, len(self.I)
        z = nuts.pack_state(X, s, tau)
        e, d = eps.reshape(-1), da.reshape(-1, 4)
        ids = torch.arange(cid0, cid0 + B * R, dtype=torch.int64, device=z.device)
        eng = nuts.FusedLeafEngine(self.prob, R)
        LB = torch.as_tensor(self.sigma_sqs_LB, dtype=torch.float64, device=z.device)[:, None]
        burn = nuts.nuts_run_(z, e, d, None, n_iter=num_burnin_steps, num_adapt=num_adapt, seed=seed, chain_ids=ids,
                              max_tree_depth=max_tree_depth, leaf_engine=eng, cached_target=cached_target)
        th = torch.empty((num_results, B, R, P), dtype=torch.float64, device=z.device)
        sg = torch.empty((num_results, B, R, D), dtype=torch.float64, device=z.device)
        Xsum = torch.zeros((B, R, n * D), dtype=torch.float64, device=z.device)
        Xsq = torch.zeros_like(Xsum)

        def on_sample(it, zz, info):
            zz = zz.view(B, R, -1)
            th[it] = torch.nn.functional.softplus(zz[..., n * D + D:])                            # :419
            sg[it] = torch.nn.functional.softplus(zz[..., n * D:n * D + D]) + LB                  # :418
            if keep_X_mean:
                Xsum.add_(zz[..., :n * D]); Xsq.addcmul_(zz[..., :n * D], zz[..., :n * D])

        out = nuts.nuts_run_(z, e, d, None, n_iter=num_results, iter0=num_burnin_steps, num_adapt=num_adapt, seed=seed,
                             chain_ids=ids, max_tree_depth=max_tree_depth, on_sample=on_sample, leaf_engine=eng,
                             cached_target=cached_target)
        if gather and self.world > 1:
            sizes = parallel.shard_sizes(self.B_total, self.world)
            th = parallel.gather_samples(th, dataset_dim=1, sizes=sizes)
            sg = parallel.gather_samples(sg, dataset_dim=1, sizes=sizes)
        res = {"thetas_samps": th.permute(1, 2, 0, 3).cpu().numpy(), "sigma_sqs_samps": sg.permute(1, 2, 0, 3).cpu().numpy(),
               "accept_prob": out["accept_prob"].mean(dim=0).view(B, R).cpu().numpy(),
               "leapfrogs_taken": out["n_leapfrog"].view(-1, B, R).cpu().numpy(),
               "step_size": e.view(B, R).cpu().numpy(),
               "burnin_accept_prob": burn["accept_prob"].mean(dim=0).view(B, R).cpu().numpy(),
               "phi1s": self.phi1s, "phi2s": self.phi2s, "thetas_init": self.thetas_init, "I": self.I,
               "dataset_range": (self.lo, self.hi)}
        if keep_X_mean:
            mean = Xsum.sum(dim=1) / (R * num_results)
            var = Xsq.sum(dim=1) / (R * num_results) - mean ** 2
            res["X_mean"] = mean.view(B, n, D).cpu().numpy()
            res["X_sd"] = var.clamp_min(0).sqrt().view(B, n, D).cpu().numpy()
        return res
