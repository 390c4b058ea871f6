"""CPU restatement of the reference's initial-fit stages.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  torch-CPU FP64 with autograd standing in for TF
autodiff, op for op along the cited lines of ``/root/reference/magi_v2.py``:

  * ``fit_kernel_hparams``   -- ``_fit_kernel_hparams`` (:538-691): Fourier prior (:549-565), TFP
    ``GeneralizedMatern`` + ``GaussianProcess`` marginal likelihood (:574-598), truncated-normal priors
    (:610-628), softplus ``TransformedVariable``s (:631-642), tf_keras Adam (:654-678).
  * ``fit_thetas_init``      -- theta initialisation when every component is observed (:132-179),
    INCLUDING the reference's ``tf.reshape`` of the [n,D] right-hand side to (D,n,1) (:155-156; a reshape,
    not a transpose) unless ``layout="transpose"`` is asked for.
  * ``fit_unobserved``       -- joint gradient matching of (X_unobs, thetas) (:182-250).
  * ``cv_cubic_smoother``    -- (:695-770), including the knot-count slip at :747-767.
  * ``initial_fit``          -- the whole of ``MAGI_v2.initial_fit`` (:82-277) composed from the above and
    ``oracle.magi_oracle``.

Parity status: TFP / tf_keras are not installable here, so everything that restates them
(``GeneralizedMatern``, ``GaussianProcess(jitter=1e-6)``, ``TruncatedNormal.log_prob``,
``JointDistributionNamed.log_prob`` broadcasting, ``tf_keras.optimizers.Adam``) is "parity unpinned";
``cv_cubic_smoother`` is pinned against the genuine reference method (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
from scipy.special import gamma as _gamma
from scipy.special import kv as _kv

from . import magi_oracle as mo

NU = 2.01
JITTER = 1e-6            # tfd.GaussianProcess default jitter (deprecation warning at output.log:15-17)


# --------------------------------------------------------------------------------------
# tf_keras.optimizers.Adam (requirements.txt:7, tf-keras==2.17.0), defaults beta_1 0.9, beta_2 0.999, epsilon 1e-7
# --------------------------------------------------------------------------------------


class KerasAdam:
    """update_step of tf_keras' Adam: alpha = lr sqrt(1 - b2^t) / (1 - b1^t); m += (g - m)(1 - b1);
    v += (g^2 - v)(1 - b2); var -= m alpha / (sqrt(v) + eps)."""

    def __init__(self, params, lr=0.01, b1=0.9, b2=0.999, eps=1e-7):
        import torch
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def apply_gradients(self, grads):
        import torch
        self.t += 1
        alpha = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                m.add_((g - m) * (1.0 - self.b1))
                v.add_((g * g - v) * (1.0 - self.b2))
                p.sub_(m * alpha / (v.sqrt() + self.eps))


# --------------------------------------------------------------------------------------
# f1: GP hyper-parameter fit (:538-691)
# --------------------------------------------------------------------------------------


def fourier_prior(X_filled: np.ndarray):
    """magi_v2.py:549-565, one component at a time as the reference loops.  X_filled [n, D]."""
    mu_phi2s, sd_phi2s = [], []
    for d in range(X_filled.shape[1]):
        z = np.fft.fft(X_filled[:, d]); zmod = np.abs(z)                                   # :552
        zmod_effective = zmod[1:(len(zmod) - 1) // 2 + 1]; zmod_effective_sq = zmod_effective ** 2
        idxs = np.linspace(1, len(zmod_effective), len(zmod_effective))
        freq = np.sum(idxs * zmod_effective_sq) / np.sum(zmod_effective_sq)                # :555
        mu_phi2 = 0.5 / freq; sd_phi2 = (1 - mu_phi2) / 3                                  # :556
        mu_phi2s.append(mu_phi2); sd_phi2s.append(sd_phi2)
    return np.array(mu_phi2s), np.array(sd_phi2s)


def _matern_corr_fn():
    """rho(z) = 2^{1-nu}/Gamma(nu) z^nu K_nu(z), rho(0) = 1 (tfk.GeneralizedMatern with amplitude 1,
    z = sqrt(2 nu) r / length_scale); d rho/dz = -2^{1-nu}/Gamma(nu) z^nu K_{nu-1}(z)."""
    import torch

    c = 2.0 ** (1.0 - NU) / _gamma(NU)

    class MaternCorr(torch.autograd.Function):
        @staticmethod
        def forward(ctx, z):
            zn = z.detach().numpy()
            pos = zn > 0
            zs = np.where(pos, zn, 1.0)
            val = np.where(pos, c * zs ** NU * _kv(NU, zs), 1.0)
            der = np.where(pos, -c * zs ** NU * _kv(NU - 1.0, zs), 0.0)
            ctx.save_for_backward(torch.from_numpy(der))
            return torch.from_numpy(val)

        @staticmethod
        def backward(ctx, g):
            (der,) = ctx.saved_tensors
            return g * der

    return MaternCorr.apply


def _softplus_inverse(a):
    return np.log(np.expm1(a))


def _truncnorm_log_prob(x, loc, scale, low):
    """tfd.TruncatedNormal(loc, scale, low, high=inf).log_prob(x) (:611-627).  torch tensors [D]."""
    import torch
    nrm = torch.distributions.Normal(0.0, 1.0)
    log_z = torch.log(1.0 - nrm.cdf((low - loc) / scale))
    lp = -0.5 * ((x - loc) / scale) ** 2 - 0.5 * math.log(2.0 * math.pi) - torch.log(scale) - log_z
    return torch.where(x >= low, lp, torch.full_like(lp, -math.inf))


class HparamObjective:
    """``gpjm.log_prob`` of magi_v2.py:610-653 as a function of the three softplus pre-activations."""

    def __init__(self, I: np.ndarray, X_filled: np.ndarray):
        import torch
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
        self.n, self.D = X_filled.shape
        D = self.D
        self.mu_phi2s, self.sd_phi2s = fourier_prior(X_filled)
        self.mu_ds = T(X_filled.mean(axis=0))                                              # :559
        sd = X_filled.std(axis=0)
        t = np.asarray(I, dtype=np.float64).reshape(-1)
        self.r = T(np.abs(t[:, None] - t[None, :]))
        self.x = T(X_filled.T)                                                             # X_filled_bcst, :645
        self.eye = torch.eye(self.n, dtype=torch.float64)
        self.loc = dict(phi1=T(np.full(D, 1e-4)), sig=T((sd * 0.1) ** 2), phi2=T(self.mu_phi2s))
        self.scale = dict(phi1=T(np.full(D, 1000.0 * np.sqrt(D))), sig=T(np.full(D, 1000.0 * np.sqrt(D))),
                          phi2=T(self.sd_phi2s * np.sqrt(D)))
        self.low = T(np.full(D, 1e-6))
        self.init = dict(phi1=sd ** 2, phi2=self.mu_phi2s, sig=(sd * 0.1) ** 2)            # :631-642
        self.corr = _matern_corr_fn()

    def initial_variables(self):
        import torch
        return [torch.tensor(_softplus_inverse(self.init[k]), dtype=torch.float64, requires_grad=True)
                for k in ("phi1", "phi2", "sig")]                                          # order of :645

    def gp_log_prob(self, phi1, sig, phi2):
        """tfd.GaussianProcess(GeneralizedMatern(2.01, sqrt(phi1), phi2), I, mean_fn, sigma^2).log_prob, [D]."""
        import torch
        z = math.sqrt(2.0 * NU) * self.r[None] / phi2[:, None, None]
        S = phi1[:, None, None] * self.corr(z) + (sig + JITTER)[:, None, None] * self.eye
        L = torch.linalg.cholesky(S)
        xc = (self.x - self.mu_ds[:, None])[..., None]
        a = torch.linalg.solve_triangular(L, xc, upper=False)[..., 0]
        return (-0.5 * (a * a).sum(-1) - torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
                - 0.5 * self.n * math.log(2.0 * math.pi))

    def log_prob(self, v_phi1, v_phi2, v_sig):
        """The [D, D] array ``gpjm.log_prob`` returns: priors are [D], the GP has batch shape [D, 1] (:578-597),
        and JointDistributionNamed adds them with broadcasting (the author's note at :603-607)."""
        import torch
        sp = torch.nn.functional.softplus
        phi1, phi2, sig = sp(v_phi1), sp(v_phi2), sp(v_sig)
        pri = (_truncnorm_log_prob(phi1, self.loc["phi1"], self.scale["phi1"], self.low)
               + _truncnorm_log_prob(sig, self.loc["sig"], self.scale["sig"], self.low)
               + _truncnorm_log_prob(phi2, self.loc["phi2"], self.scale["phi2"], self.low))
        return pri[None, :] + self.gp_log_prob(phi1, sig, phi2)[:, None]

    def loss_and_grads(self, variables):
        """train_model (:657-664): loss = -log_prob ([D, D]); tape.gradient of a non-scalar sums it."""
        import torch
        loss = -self.log_prob(*variables)
        grads = torch.autograd.grad(loss.sum(), variables)
        return loss.detach(), grads


def fit_kernel_hparams(I: np.ndarray, X_filled: np.ndarray, num_iters: int = 1000, lr: float = 0.01,
                       trace: Optional[list] = None):
    """magi_v2.py:538-691.  Returns {"phi1s", "phi2s", "sigma_sqs"} each [D]."""
    import torch
    obj = HparamObjective(I, X_filled)
    variables = obj.initial_variables()
    opt = KerasAdam(variables, lr=lr)
    for _ in range(num_iters):
        loss, grads = obj.loss_and_grads(variables)
        opt.apply_gradients(grads)
        if trace is not None:
            trace.append([v.detach().numpy().copy() for v in variables])
    sp = torch.nn.functional.softplus
    phi1, phi2, sig = (sp(v).detach().numpy().copy() for v in variables)
    return {"phi1s": phi1, "phi2s": phi2, "sigma_sqs": sig}


# --------------------------------------------------------------------------------------
# f3: theta initialisation (:132-179) and unobserved components (:182-250)
# --------------------------------------------------------------------------------------


def theta_objective(thetas, I, Xhat_init, mu_ds, m_ds, K_d_invs, f_vec: Callable, layout: str = "reference"):
    """magi_v2.py:150-158.  torch tensors.  layout "reference": tf.reshape of the [n, D] right-hand side to
    (D, n, 1) as written at :155-156; "transpose": the (D, n, 1) layout ``unnormalized_log_prob`` uses (:335)."""
    import torch
    n, D = Xhat_init.shape
    X_cent = torch.reshape(Xhat_init - mu_ds, (n, 1, D))                                   # :139-141
    m_ds_prod_X_cent = m_ds @ X_cent.permute(2, 0, 1)                                      # :142
    f = f_vec(I, Xhat_init, thetas)
    if layout == "reference":
        f_vals = torch.reshape(f, (D, n, 1))                                               # :155-156
    else:
        f_vals = f.T.reshape(D, n, 1)
    toNorm = f_vals - m_ds_prod_X_cent                                                     # :157
    return torch.sum(toNorm.permute(0, 2, 1) @ (K_d_invs @ toNorm))                        # :158


def fit_thetas_init(I, Xhat_init, mu_ds, m_ds, K_d_invs, f_vec: Callable, D_thetas: int, num_iters: int = 10000,
                    lr: float = 0.01, layout: str = "reference"):
    """magi_v2.py:132-179: Adam (lr 0.01, 10 000 steps) from theta = 1 on the t2-only objective, evaluated at the
    linearly interpolated Xhat_init with the UN-banded matrices (band_part runs afterwards, :271-274)."""
    import torch
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
    I_t, X_t, mu_t, m_t, K_t = T(np.asarray(I).reshape(-1, 1)), T(Xhat_init), T(mu_ds), T(m_ds), T(K_d_invs)
    thetas = torch.ones(D_thetas, dtype=torch.float64, requires_grad=True)                 # :136
    opt = KerasAdam([thetas], lr=lr)                                                       # :161
    for _ in range(num_iters):
        loss = theta_objective(thetas, I_t, X_t, mu_t, m_t, K_t, f_vec, layout)
        opt.apply_gradients(torch.autograd.grad(loss, [thetas]))
    return thetas.detach().numpy().copy()


def fit_unobserved(I, X_smoothed_obs, X_interp_obs, observed_components, unobserved_components, f_vec: Callable,
                   D_thetas: int, X_unobs_start: np.ndarray, num_iters: int = 10000, lr: float = 0.01):
    """magi_v2.py:196-250 given the random start ``X_unobs_start`` (the reference draws it unseeded from
    N(mean, sd) of the interpolated observed data, :220-227).  Returns (X_unobs, thetas)."""
    import torch
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
    order = torch.as_tensor(np.argsort(np.concatenate([observed_components, unobserved_components])))  # :50
    Xo, I_t = T(X_smoothed_obs), T(np.asarray(I).reshape(-1, 1))
    Xu = T(X_unobs_start).clone().requires_grad_(True)
    th = torch.ones(D_thetas, dtype=torch.float64, requires_grad=True)                     # :228
    dx2 = 2.0 * float(I_t[1, 0] - I_t[0, 0])                                               # :213
    opt = KerasAdam([Xu, th], lr=lr)                                                       # :231
    for _ in range(num_iters):
        X_full = torch.cat([Xo, Xu], dim=1)[:, order]                                      # :202-203
        f_vals = f_vec(I_t, X_full, th)                                                    # :206
        f_diff = (X_full[2:, :] - X_full[:-2, :]) / dx2                                    # :213
        loss = torch.sum((f_vals[1:-1] - f_diff) ** 2)                                     # :216
        opt.apply_gradients(torch.autograd.grad(loss, [Xu, th]))
    return Xu.detach().numpy().copy(), th.detach().numpy().copy()


# --------------------------------------------------------------------------------------
# spline smoother (:695-770)
# --------------------------------------------------------------------------------------


def single_cv_cubic_smoother(I, x):
    """magi_v2.py:708-770.  The cross-validation picks ``optimal_knot_num`` (:747) but the final fit uses the
    loop variable ``knot_num``, i.e. the LARGEST knot count (:750-767); reproduced as written."""
    from scipy.interpolate import splev, splrep
    from sklearn.model_selection import KFold
    I = np.asarray(I).flatten()
    if I.shape[0] < 10:
        return x
    kf = KFold(n_splits=5, shuffle=True, random_state=1)                                   # :715
    knot_nums = np.arange(0, (I.shape[0] // 10) + 1)                                       # :718
    split_errs = []
    for train_idx, val_idx in kf.split(np.arange(I.shape[0])):
        knot_errs = []
        for knot_num in knot_nums:
            knots = np.array([]) if knot_num == 0 else np.linspace(I[0], I[-1], knot_num + 2)[1:-1]
            tck = splrep(I[train_idx], x[train_idx], t=knots, s=0)
            preds = splev(I[val_idx], tck)
            knot_errs.append(((preds - x[val_idx]) ** 2).mean())
        split_errs.append(knot_errs)
    _optimal_knot_num = knot_nums[np.array(split_errs).mean(axis=0).argmin()]              # :747 (unused, as there)
    knots = np.array([]) if knot_num == 0 else np.linspace(I[0], I[-1], knot_num + 2)[1:-1]  # :750-753
    tck = splrep(I, x, t=knots, s=0)
    return splev(I, tck)


def cv_cubic_smoother(I, X_filled):
    """magi_v2.py:695-705."""
    I = np.asarray(I).flatten()
    if I.shape[0] < 10:
        return X_filled
    return np.stack([single_cv_cubic_smoother(I, X_filled[:, i]) for i in range(X_filled.shape[1])], axis=1)


# --------------------------------------------------------------------------------------
# MAGI_v2.initial_fit (:82-277), every component observed
# --------------------------------------------------------------------------------------


def initial_fit(ts_obs, X_obs, discretization: int, bandsize: Optional[int], f_vec: Callable, D_thetas: int,
                hparams: Optional[dict] = None, hparam_iters: int = 1000, theta_iters: int = 10000,
                theta_layout: str = "reference"):
    """The reference's ``initial_fit`` for fully observed systems, returning what ``predict`` reads:
    dict(I, phi1s, phi2s, sigma_sqs_init, Xhat_init, thetas_init, constants=PosteriorConstants)."""
    X_obs = np.asarray(X_obs, dtype=np.float64)
    I, X_obs_discret = mo.discretize(ts_obs, X_obs, discretization)                        # :85
    X_interp = mo.linear_interpolate(X_obs_discret)                                        # :105
    if hparams is None:
        hparams = fit_kernel_hparams(I, X_interp, num_iters=hparam_iters)                  # :106
    phi1s, phi2s = np.asarray(hparams["phi1s"]), np.asarray(hparams["phi2s"])
    mu_ds = X_interp.mean(axis=0)                                                          # :114
    dense = mo.kernel_matrices(I, phi1s, phi2s, None)                                      # :122-128
    thetas_init = fit_thetas_init(I, X_interp, mu_ds, dense[1], dense[2], f_vec, D_thetas, theta_iters,
                                  layout=theta_layout)                                     # :133-179
    banded = tuple(mo.band_part(A, bandsize) for A in dense)                               # :271-274
    Xhat_init = cv_cubic_smoother(I, X_interp)                                             # :277
    c = mo.make_constants(ts_obs, X_obs, discretization, phi1s, phi2s, bandsize, f_vec, Xhat_init=Xhat_init,
                          matrices=banded)
    return dict(I=I, phi1s=phi1s, phi2s=phi2s, sigma_sqs_init=np.asarray(hparams["sigma_sqs"]),
                Xhat_init=Xhat_init, thetas_init=thetas_init, constants=c)
