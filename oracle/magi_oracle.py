"""CPU restatement of the reference's posterior-evaluation path (numpy / torch-CPU FP64).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py for who may import this and
for the parity-pinning status of each function.  Every function cites the
reference lines (``/root/reference/magi_v2.py`` unless stated) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np
from scipy.special import gamma as _gamma
from scipy.special import kvp as _kvp

# --------------------------------------------------------------------------------------
# ODE right-hand sides (the user-supplied ``f_vec(t, X[n,D], thetas[P]) -> [n,D]``)
# --------------------------------------------------------------------------------------


def _cat(cols, like):
    if isinstance(like, np.ndarray):
        return np.concatenate(cols, axis=1)
    import torch

    return torch.cat(cols, dim=1)


def f_seir3(t, X, th):
    """vignette.ipynb:68-79 -- (E, I, R), S = 1 - E - I - R implicit; theta = (beta, gamma, sigma)."""
    S = 1.0 - X.sum(axis=1).reshape(-1, 1)
    return _cat([(th[0] * S * X[:, 1:2]) - (th[2] * X[:, 0:1]),
                 (th[2] * X[:, 0:1]) - (th[1] * X[:, 1:2]),
                 (th[1] * X[:, 1:2])], X)


def f_seir4(t, X, th):
    """SEIR with S explicit (BASELINE.json throughput shape D=4):
    dS=-bSI, dE=bSI-sE, dI=sE-gI, dR=gI; theta = (beta, gamma, sigma)."""
    S, E, I_, R = X[:, 0:1], X[:, 1:2], X[:, 2:3], X[:, 3:4]
    return _cat([-th[0] * S * I_, th[0] * S * I_ - th[2] * E, th[2] * E - th[1] * I_, th[1] * I_], X)


def f_sirw(t, X, th):
    """test_magi_script.py:19-45 -- (S, I, R, W); theta = (beta, phi, xi, chi, kappa)."""
    S, I_, R, W = X[:, 0:1], X[:, 1:2], X[:, 2:3], X[:, 3:4]
    beta, phi, xi, chi, kappa = th[0], th[1], th[2], th[3], th[4]
    return _cat([-beta * S * I_ + kappa * W,
                 beta * S * I_ - phi * I_,
                 phi * I_ - xi * R + chi * I_ * W,
                 xi * R - chi * I_ * W - kappa * W], X)


def f_lorenz96(t, X, th):
    """Lorenz-96: dx_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F, P=1."""
    D = X.shape[1]
    cols = []
    for i in range(D):
        ip1, im1, im2 = (i + 1) % D, (i - 1) % D, (i - 2) % D
        cols.append((X[:, ip1:ip1 + 1] - X[:, im2:im2 + 1]) * X[:, im1:im1 + 1] - X[:, i:i + 1] + th[0])
    return _cat(cols, X)


@dataclass(frozen=True)
class OdeModel:
    name: str
    model_id: int
    D: int
    P: int
    f_vec: Callable


MODELS: Dict[str, OdeModel] = {
    "seir3": OdeModel("seir3", 0, 3, 3, f_seir3),
    "seir4": OdeModel("seir4", 1, 4, 3, f_seir4),
    "sirw": OdeModel("sirw", 2, 4, 5, f_sirw),
    "lorenz96": OdeModel("lorenz96", 3, 10, 1, f_lorenz96),
}


# --------------------------------------------------------------------------------------
# L1: covariance build, inversion, banding
# --------------------------------------------------------------------------------------


def matern_blocks(I, phi1, phi2, v=2.01):
    """magi_v2.py:781-815 restated: (Kappa, p_Kappa, Kappa_pp) on the grid I (n,) or (n,1)."""
    I = np.asarray(I, dtype=np.float64)
    s = np.tile(A=I.reshape(-1, 1), reps=I.size)          # :781
    t = s.T
    l = np.abs(s - t)                                        # :784
    u = np.sqrt(2 * v) * l / phi2
    np.fill_diagonal(u, np.nan)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        Bv0, Bv1, Bv2 = _kvp(v, u, 0), _kvp(v, u, 1), _kvp(v, u, 2)   # :787
        Kappa = (phi1 / _gamma(v)) * (2 ** (1 - (v / 2))) * ((np.sqrt(v) / phi2) ** v)   # :790
        Kappa = Kappa * Bv0
        Kappa = Kappa * (l ** v)
        np.fill_diagonal(Kappa, phi1)                        # :795
        p_Kappa = (2 ** (1 - (v / 2)))                       # :798
        p_Kappa = p_Kappa * (phi1 * ((u / np.sqrt(2)) ** v))   # `a *= b * c` is a * (b * c)
        p_Kappa = p_Kappa * ((u * phi2 * Bv1) + (v * phi2 * Bv0))
        p_Kappa = p_Kappa / (phi2 * (s - t) * _gamma(v))
        np.fill_diagonal(p_Kappa, 0.0)                       # :802
        Kappa_pp = 2 * np.sqrt(2) * (v ** 1.5) * phi2 * l * Bv1    # :808
        Kappa_pp = Kappa_pp + (((v ** 2) * (phi2 ** 2)) - (v * (phi2 ** 2))) * Bv0
        Kappa_pp = Kappa_pp + ((2 * v * (s ** 2)) - (4 * v * s * t) + (2 * v * (t ** 2))) * Bv2
        Kappa_pp = Kappa_pp * (-1.0 * (2 ** (1 - (v / 2))) * phi1 * ((u / np.sqrt(2)) ** v))
        Kappa_pp = Kappa_pp / ((phi2 ** 2) * (l ** 2) * _gamma(v))
        np.fill_diagonal(Kappa_pp, v * phi1 / ((phi2 ** 2) * (v - 1)))   # :815
    return Kappa, p_Kappa, Kappa_pp


def matern_blocks_roundoff_scale(I, phi1, phi2, v=2.01):
    """Per-entry size of the terms the reference's formulas for p_Kappa (:798-801) and Kappa_pp
    (:808-812) add and subtract.  k * eps * scale bounds the REFERENCE's own rounding error: its
    Kappa_pp forms 2v(s-t)^2 as 2v s^2 - 4v s t + 2v t^2 (:810), which cancels catastrophically for
    |s-t| << |s| (relative error ~ eps * s^2 / l^2, e.g. 5e-10 at t = 3.3, l = 0.025).  Tests compare a
    more accurate evaluation against the reference within this bound instead of a flat tolerance."""
    I = np.asarray(I, dtype=np.float64)
    s = np.tile(A=I.reshape(-1, 1), reps=I.size)
    t = s.T
    l = np.abs(s - t)
    u = np.sqrt(2 * v) * l / phi2
    np.fill_diagonal(u, np.nan)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        Bv0, Bv1, Bv2 = np.abs(_kvp(v, u, 0)), np.abs(_kvp(v, u, 1)), np.abs(_kvp(v, u, 2))
        pref = (2 ** (1 - (v / 2))) * phi1 * ((u / np.sqrt(2)) ** v)
        M_pK = pref * ((u * phi2 * Bv1) + (v * phi2 * Bv0)) / (phi2 * l * _gamma(v))
        M_Kpp = (2 * np.sqrt(2) * (v ** 1.5) * phi2 * l * Bv1 + ((v ** 2) * (phi2 ** 2) + v * (phi2 ** 2)) * Bv0
                 + ((2 * v * (s ** 2)) + (4 * v * np.abs(s * t)) + (2 * v * (t ** 2))) * Bv2)
        M_Kpp = M_Kpp * pref / ((phi2 ** 2) * (l ** 2) * _gamma(v))
    np.fill_diagonal(M_pK, 0.0)
    np.fill_diagonal(M_Kpp, 0.0)
    return np.nan_to_num(M_pK), np.nan_to_num(M_Kpp)


def build_matrices(I, phi1, phi2, v=2.01):
    """magi_v2.py:774-823 restated: returns (C_d, m_d, K_d)."""
    Kappa, p_Kappa, Kappa_pp = matern_blocks(I, phi1, phi2, v)
    Kappa_p = p_Kappa * -1                                   # :805
    C_d, Kappa_inv = Kappa.copy(), np.linalg.pinv(Kappa)     # :818
    m_d = p_Kappa @ Kappa_inv                                # :819
    K_d = Kappa_pp - (p_Kappa @ Kappa_inv @ Kappa_p)         # :820
    return C_d, m_d, K_d


def tf_pinv(A):
    """Stand-in for ``tf.linalg.pinv(a)`` (magi_v2.py:126,128): SVD pseudo-inverse with TF's
    default cutoff rcond = 10 * max(rows, cols) * eps."""
    A = np.asarray(A, dtype=np.float64)
    return np.linalg.pinv(A, rcond=10.0 * max(A.shape) * np.finfo(np.float64).eps)


def band_part(A, b: Optional[int]):
    """``tf.linalg.band_part(A, b, b)`` on the last two axes (magi_v2.py:271-274)."""
    if b is None:
        return np.array(A, copy=True)
    n = A.shape[-1]
    i, j = np.indices((n, n))
    return np.where(np.abs(i - j) <= b, A, 0.0)


def kernel_matrices(I, phi1s, phi2s, bandsize: Optional[int], v=2.01):
    """magi_v2.py:116-128 + :271-274: per component (C_d^-1, m_d, K_d^-1), banded, each [D,n,n]."""
    D = len(phi1s)
    n = np.asarray(I).size
    Cinv, m, Kinv = (np.zeros((D, n, n)) for _ in range(3))
    for d in range(D):
        C_d, m_d, K_d = build_matrices(I, phi1s[d], phi2s[d], v)
        Cinv[d] = tf_pinv(C_d)
        m[d] = m_d
        Kinv[d] = tf_pinv(K_d)
    return band_part(Cinv, bandsize), band_part(m, bandsize), band_part(Kinv, bandsize)


# --------------------------------------------------------------------------------------
# L0 helpers that define the constants of the log-posterior
# --------------------------------------------------------------------------------------


def discretize(ts_obs, X_obs, discretization):
    """magi_v2.py:475-498."""
    ts_obs = np.asarray(ts_obs, dtype=np.float64).flatten()
    assert ts_obs.shape[0] == X_obs.shape[0], \
        "Please make sure there are equal numbers of observations in ts_obs and X_obs."
    N, D = X_obs.shape
    step = 2 ** discretization
    N_discret = step * (N - 1) + 1
    I = np.full((N_discret,), np.nan)
    X_obs_discret = np.full((N_discret, D), np.nan)
    I[::step] = ts_obs
    indices = np.arange(len(I))
    I = np.interp(x=indices, xp=indices[~np.isnan(I)], fp=I[~np.isnan(I)])
    X_obs_discret[::step] = X_obs
    return I.reshape(-1, 1), X_obs_discret


def linear_interpolate(X_partial):
    """magi_v2.py:509-527."""
    N_partial, D_partial = X_partial.shape
    X_interp = X_partial.copy()
    indices = np.arange(N_partial)
    for d in range(D_partial):
        if np.any(np.isnan(X_interp[:, d])):
            ok = ~np.isnan(X_partial[:, d])
            X_interp[:, d] = np.interp(x=indices, xp=indices[ok], fp=X_partial[ok, d])
    return X_interp


@dataclass
class PosteriorConstants:
    """Everything ``unnormalized_log_prob`` closes over (magi_v2.py:294-300)."""
    I: np.ndarray            # [n,1]
    mu_ds: np.ndarray        # [D]
    C_d_invs: np.ndarray     # [D,n,n]
    m_ds: np.ndarray         # [D,n,n]
    K_d_invs: np.ndarray     # [D,n,n]
    N_ds: np.ndarray         # [D]   non-NaN raw observation counts (:53)
    not_nan_idxs: np.ndarray  # [#obs] flattened row-major indices into [n,D] (:96)
    not_nan_cols: np.ndarray  # [#obs] = idx % D (:97)
    y_tau_ds_observed: np.ndarray  # [#obs] (:100)
    beta: float              # D*n / sum(N_ds) (:89)
    sigma_sqs_LB: np.ndarray  # [D] (:299-300)
    f_vec: Callable

    @property
    def n(self):
        return self.I.shape[0]

    @property
    def D(self):
        return self.mu_ds.shape[0]

    def dense_y_mask(self):
        """The same information as (not_nan_idxs, y_tau_ds_observed) in the dense layout the
        C-ABI takes: y[n,D] (0 where unobserved) and mask[n,D] uint8."""
        n, D = self.n, self.D
        y = np.zeros(n * D)
        mask = np.zeros(n * D, dtype=np.uint8)
        y[self.not_nan_idxs] = self.y_tau_ds_observed
        mask[self.not_nan_idxs] = 1
        return y.reshape(n, D), mask.reshape(n, D)


def make_constants(ts_obs, X_obs, discretization, phi1s, phi2s, bandsize, f_vec,
                   Xhat_init=None, sigma_sqs_LB=None, v=2.01, matrices=None) -> PosteriorConstants:
    """The bookkeeping of ``__init__`` (:42-53) and ``initial_fit`` (:85-128, :271-274) that feeds
    the log-posterior, given kernel hyper-parameters (the GP hyper-parameter fit itself,
    :538-691, is outside the hot path)."""
    X_obs = np.asarray(X_obs, dtype=np.float64)
    N_ds = (~np.isnan(X_obs)).sum(axis=0)                                  # :53
    I, X_obs_discret = discretize(ts_obs, X_obs, discretization)            # :85
    n, D = X_obs_discret.shape
    beta = float((D * n) / N_ds.sum())                                      # :89
    not_nan_idxs = np.where(~np.isnan(X_obs_discret).flatten())[0]          # :96
    not_nan_cols = not_nan_idxs % D                                         # :97
    y_obs = X_obs_discret.reshape(-1)[not_nan_idxs]                         # :100
    X_interp = linear_interpolate(X_obs_discret)                            # :105
    mu_ds = X_interp.mean(axis=0)                                           # :114
    if matrices is None:
        matrices = kernel_matrices(I, phi1s, phi2s, bandsize, v)            # :122-128, :271-274
    Cinv, m, Kinv = matrices
    if Xhat_init is None:
        Xhat_init = X_interp
    if sigma_sqs_LB is None:
        sigma_sqs_LB = (Xhat_init.std(axis=0) * 0.01) ** 2                  # :299-300
    return PosteriorConstants(I=I, mu_ds=mu_ds, C_d_invs=Cinv, m_ds=m, K_d_invs=Kinv, N_ds=N_ds,
                              not_nan_idxs=not_nan_idxs, not_nan_cols=not_nan_cols,
                              y_tau_ds_observed=y_obs, beta=beta,
                              sigma_sqs_LB=np.asarray(sigma_sqs_LB, dtype=np.float64), f_vec=f_vec)


# --------------------------------------------------------------------------------------
# L2: the log-posterior (magi_v2.py:308-348), three ways
# --------------------------------------------------------------------------------------


def log_posterior(X, sigma_sqs_pre, thetas_pre, beta_temp, c: PosteriorConstants):
    """numpy op-for-op restatement of ``unnormalized_log_prob`` (magi_v2.py:308-348)."""
    sigma_sqs = np.log(1.0 + np.exp(sigma_sqs_pre)) + c.sigma_sqs_LB                    # :318
    thetas = np.log(1.0 + np.exp(thetas_pre))                                           # :319
    log_jacobian_sigma_sqs = np.sum(sigma_sqs_pre - np.log(1.0 + np.exp(sigma_sqs_pre)))  # :322
    log_jacobian_thetas = np.sum(thetas_pre - np.log(1.0 + np.exp(thetas_pre)))         # :323
    X_cent = np.reshape(X - c.mu_ds, (X.shape[0], 1, X.shape[1]))                       # :329
    Xc_rev = np.transpose(X_cent)                    # tf.transpose w/o perm reverses axes -> (D,1,n)
    Xc_col = np.transpose(X_cent, (2, 0, 1))         # (D,n,1)
    t1 = np.sum((Xc_rev @ c.C_d_invs) @ Xc_col)                                         # :332
    f_vals = np.transpose(c.f_vec(c.I, X, thetas)[:, None], (2, 0, 1))                  # :335
    toNorm = f_vals - (c.m_ds @ Xc_col)                                                 # :336
    t2 = np.sum(np.transpose(toNorm, (0, 2, 1)) @ (c.K_d_invs @ toNorm))                # :337
    t3 = np.sum(c.N_ds * np.log(2.0 * np.pi * sigma_sqs))                               # :340
    X_observed = np.reshape(X, [-1])[c.not_nan_idxs]                                    # :343
    t4 = np.sum(np.square(X_observed - c.y_tau_ds_observed) * (1.0 / sigma_sqs)[c.not_nan_cols])  # :344
    return beta_temp * (-0.5 * (((1.0 / c.beta) * (t1 + t2)) + (t3 + t4))
                        + log_jacobian_sigma_sqs + log_jacobian_thetas)                 # :348


def log_posterior_and_grad_autograd(X, sigma_sqs_pre, thetas_pre, beta_temp, c: PosteriorConstants):
    """torch-CPU FP64 op-for-op restatement of magi_v2.py:308-348; the gradient w.r.t. the three
    state parts comes from reverse-mode autodiff, as TFP's leapfrog obtains it in the reference
    (value_and_gradient of target_log_prob_fn, :362, :867)."""
    import torch

    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    X_t = T(X).clone().requires_grad_(True)
    s_t = T(sigma_sqs_pre).clone().requires_grad_(True)
    th_t = T(thetas_pre).clone().requires_grad_(True)
    mu, Cinv, m, Kinv = T(c.mu_ds), T(c.C_d_invs), T(c.m_ds), T(c.K_d_invs)
    LB, N_ds, y = T(c.sigma_sqs_LB), T(c.N_ds), T(c.y_tau_ds_observed)
    idx = torch.as_tensor(c.not_nan_idxs, dtype=torch.int64)
    cols = torch.as_tensor(c.not_nan_cols, dtype=torch.int64)
    I_t = T(c.I)
    bt = float(beta_temp)                                                               # :326 stop_gradient
    sigma_sqs = torch.log(1.0 + torch.exp(s_t)) + LB
    thetas = torch.log(1.0 + torch.exp(th_t))
    lj_s = torch.sum(s_t - torch.log(1.0 + torch.exp(s_t)))
    lj_t = torch.sum(th_t - torch.log(1.0 + torch.exp(th_t)))
    X_cent = torch.reshape(X_t - mu, (X_t.shape[0], 1, X_t.shape[1]))
    Xc_rev = X_cent.permute(2, 1, 0)
    Xc_col = X_cent.permute(2, 0, 1)
    t1 = torch.sum((Xc_rev @ Cinv) @ Xc_col)
    f_vals = c.f_vec(I_t, X_t, thetas)[:, None].permute(2, 0, 1)
    toNorm = f_vals - (m @ Xc_col)
    t2 = torch.sum(toNorm.permute(0, 2, 1) @ (Kinv @ toNorm))
    t3 = torch.sum(N_ds * torch.log(2.0 * math.pi * sigma_sqs))
    X_observed = torch.reshape(X_t, [-1])[idx]
    t4 = torch.sum(torch.square(X_observed - y) * (1.0 / sigma_sqs)[cols])
    lp = bt * (-0.5 * (((1.0 / c.beta) * (t1 + t2)) + (t3 + t4)) + lj_s + lj_t)
    lp.backward()
    return (float(lp.detach()), X_t.grad.numpy().copy(), s_t.grad.numpy().copy(), th_t.grad.numpy().copy())


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def model_jacobians(model: str, X, th):
    """Analytic J[d',d](i) = d f_d'/d x_d and G[d',k](i) = d f_d'/d theta_k for the registry models
    (what reverse-mode autodiff of ``f_vec`` yields in the reference, magi_v2.py:335)."""
    n, D = X.shape
    P = len(th)
    J = np.zeros((n, D, D))
    G = np.zeros((n, D, P))
    if model == "seir3":
        E, I_, R = X[:, 0], X[:, 1], X[:, 2]
        S = 1.0 - E - I_ - R
        b, g, s = th
        J[:, 0, 0] = -b * I_ - s; J[:, 0, 1] = b * S - b * I_; J[:, 0, 2] = -b * I_
        J[:, 1, 0] = s; J[:, 1, 1] = -g
        J[:, 2, 1] = g
        G[:, 0, 0] = S * I_; G[:, 0, 2] = -E
        G[:, 1, 1] = -I_; G[:, 1, 2] = E
        G[:, 2, 1] = I_
    elif model == "seir4":
        S, E, I_, R = X.T
        b, g, s = th
        J[:, 0, 0] = -b * I_; J[:, 0, 2] = -b * S
        J[:, 1, 0] = b * I_; J[:, 1, 1] = -s; J[:, 1, 2] = b * S
        J[:, 2, 1] = s; J[:, 2, 2] = -g
        J[:, 3, 2] = g
        G[:, 0, 0] = -S * I_
        G[:, 1, 0] = S * I_; G[:, 1, 2] = -E
        G[:, 2, 1] = -I_; G[:, 2, 2] = E
        G[:, 3, 1] = I_
    elif model == "sirw":
        S, I_, R, W = X.T
        beta, phi, xi, chi, kappa = th
        J[:, 0, 0] = -beta * I_; J[:, 0, 1] = -beta * S; J[:, 0, 3] = kappa
        J[:, 1, 0] = beta * I_; J[:, 1, 1] = beta * S - phi
        J[:, 2, 1] = phi + chi * W; J[:, 2, 2] = -xi; J[:, 2, 3] = chi * I_
        J[:, 3, 1] = -chi * W; J[:, 3, 2] = xi; J[:, 3, 3] = -chi * I_ - kappa
        G[:, 0, 0] = -S * I_; G[:, 0, 4] = W
        G[:, 1, 0] = S * I_; G[:, 1, 1] = -I_
        G[:, 2, 1] = I_; G[:, 2, 2] = -R; G[:, 2, 3] = I_ * W
        G[:, 3, 2] = R; G[:, 3, 3] = -I_ * W; G[:, 3, 4] = -W
    elif model == "lorenz96":
        for i in range(D):
            ip1, im1, im2 = (i + 1) % D, (i - 1) % D, (i - 2) % D
            J[:, i, ip1] += X[:, im1]
            J[:, i, im2] += -X[:, im1]
            J[:, i, im1] += X[:, ip1] - X[:, im2]
            J[:, i, i] += -1.0
            G[:, i, 0] = 1.0
    else:
        raise KeyError(model)
    return J, G


def log_posterior_and_grad_analytic(X, sigma_sqs_pre, thetas_pre, beta_temp, c: PosteriorConstants,
                                    model: str):
    """Independent numpy implementation: the value as SURVEY.md A.2 and the gradient by the closed
    forms of SURVEY.md A.3 (no autodiff).  Used to cross-check the autograd restatement and as the
    vectorised CPU baseline.  Does NOT assume C^-1 / K^-1 symmetric."""
    n, D = X.shape
    s = np.asarray(sigma_sqs_pre, dtype=np.float64)
    tau = np.asarray(thetas_pre, dtype=np.float64)
    sp = lambda z: np.logaddexp(0.0, z)
    sig2 = sp(s) + c.sigma_sqs_LB
    th = sp(tau)
    xc = (X - c.mu_ds).T                                    # [D,n]
    SC = c.C_d_invs + np.transpose(c.C_d_invs, (0, 2, 1))
    SK = c.K_d_invs + np.transpose(c.K_d_invs, (0, 2, 1))
    u = np.einsum("dij,dj->di", c.C_d_invs, xc)
    t1 = np.sum(xc * u)
    f = c.f_vec(c.I, X, th).T                               # [D,n]
    r = f - np.einsum("dij,dj->di", c.m_ds, xc)
    q = np.einsum("dij,dj->di", c.K_d_invs, r)
    t2 = np.sum(r * q)
    t3 = np.sum(c.N_ds * np.log(2.0 * np.pi * sig2))
    y, mask = c.dense_y_mask()
    e = mask * (X - y)
    SSE = np.sum(e * e, axis=0)
    t4 = np.sum(SSE / sig2)
    logJ = np.sum(s - sp(s)) + np.sum(tau - sp(tau))
    lp = beta_temp * (-0.5 * ((t1 + t2) / c.beta + t3 + t4) + logJ)
    g = np.einsum("dij,dj->di", SK, r)                      # [D,n]  d t2 / d r
    J, G = model_jacobians(model, X, th)
    gX_prior = (np.einsum("dij,dj->di", SC, xc)
                + np.einsum("idc,di->ci", J, g)
                - np.einsum("dij,di->dj", c.m_ds, g))       # [D,n]
    gX = beta_temp * (-0.5) * (gX_prior.T / c.beta + 2.0 * e / sig2)
    gth = beta_temp * (-0.5 / c.beta * np.einsum("idk,di->k", G, g) * _sigmoid(tau) + (1.0 - _sigmoid(tau)))
    gs = beta_temp * (-0.5 * (c.N_ds / sig2 - SSE / sig2 ** 2) * _sigmoid(s) + (1.0 - _sigmoid(s)))
    return float(lp), gX, gs, gth


# --------------------------------------------------------------------------------------
# L3: sampler pieces (reference glue + restated TFP 0.24.0 algorithms)
# --------------------------------------------------------------------------------------


def logarithmic_temperature_schedule(step, min_temp=0.1):
    """magi_v2.py:833-835."""
    return np.maximum(1.0 / np.log(np.asarray(step, dtype=np.float64) + 2.0), min_temp)


def initial_state(Xhat_init, sigma_sqs_init, thetas_init, sigma_sqs_LB):
    """magi_v2.py:373-383: inverse-softplus initial values, -5.0 where not representable."""
    sigma_sqs_init = np.asarray(sigma_sqs_init, dtype=np.float64)
    thetas_init = np.asarray(thetas_init, dtype=np.float64)
    s0 = np.full_like(sigma_sqs_init, -5.0)
    ok = sigma_sqs_init > sigma_sqs_LB
    s0[ok] = np.log(np.exp((sigma_sqs_init - sigma_sqs_LB)[ok]) - 1.0)
    t0 = np.full_like(thetas_init, -5.0)
    ok = thetas_init > 0.0
    t0[ok] = np.log(np.exp(thetas_init[ok]) - 1.0)
    return [np.array(Xhat_init, dtype=np.float64, copy=True), s0, t0]


def pack_state(X, s, tau):
    return np.concatenate([np.ravel(X), np.ravel(s), np.ravel(tau)])


def unpack_state(z, n, D, P):
    return z[: n * D].reshape(n, D), z[n * D: n * D + D], z[n * D + D: n * D + D + P]


def leapfrog(z, p, eps, n_steps, value_and_grad):
    """TFP ``SimpleLeapfrogIntegrator`` (tensorflow-probability==0.24.0,
    mcmc/internal/leapfrog_integrator.py), identity mass:
    p_half = p + eps/2 grad(z); z' = z + eps p_half; p' = p_half + eps/2 grad(z').
    ``value_and_grad(z) -> (lp, grad)`` on the packed state.  Returns (z, p, lp, grad) and the
    trajectory of z after every step."""
    lp, g = value_and_grad(z)
    traj = []
    z = z.copy()
    p = p.copy()
    for _ in range(n_steps):
        p = p + 0.5 * eps * g
        z = z + eps * p
        lp, g = value_and_grad(z)
        p = p + 0.5 * eps * g
        traj.append(z.copy())
    return z, p, lp, g, traj


# ---- counter-based RNG shared bit-for-bit with the CUDA sampler (Philox4x32-10) ----------

_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11).  ctr: uint32 [...,4]; key: uint32 [...,2]."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c[0].astype(np.uint64)
            p1 = _PH_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _PH_W0).astype(np.uint32)
            k1 = (k1 + _PH_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def _u53(hi, lo):
    """(0,1) double from two uint32: ((hi<<21) ^ (lo>>11) as 53 bits + 0.5) * 2^-53."""
    k = (hi.astype(np.uint64) << np.uint64(21)) ^ (lo.astype(np.uint64) >> np.uint64(11))
    return (k.astype(np.float64) + 0.5) * (2.0 ** -53)


RNG_PURPOSE_MOMENTUM = 0
RNG_PURPOSE_ACCEPT = 1


def rng_normals(seed: int, chain_id: int, iteration: int, count: int):
    """``count`` standard normals for (chain, iteration): pair j -> Philox(ctr=(j, chain, iter,
    purpose=0), key=seed) -> Box-Muller (z0 = r cos, z1 = r sin).  Same counters as csrc/rng.cuh."""
    npair = (count + 1) // 2
    ctr = np.zeros((npair, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(npair, dtype=np.uint32)
    ctr[:, 1] = np.uint32(chain_id)
    ctr[:, 2] = np.uint32(iteration)
    ctr[:, 3] = np.uint32(RNG_PURPOSE_MOMENTUM)
    key = np.zeros((npair, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key)
    u1, u2 = _u53(r[:, 0], r[:, 1]), _u53(r[:, 2], r[:, 3])
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * npair)
    z[0::2] = rad * np.cos(2.0 * np.pi * u2)
    z[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return z[:count]


def rng_uniform(seed: int, chain_id: int, iteration: int):
    ctr = np.array([[0, chain_id, iteration, RNG_PURPOSE_ACCEPT]], dtype=np.uint32)
    key = np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    return float(_u53(r[:, 0], r[:, 1])[0])


@dataclass
class DualAveragingState:
    """tfp.mcmc.DualAveragingStepSizeAdaptation (0.24.0) kernel-result fields, restated
    (SURVEY.md Appendix C): target 0.75 (:366), exploration_shrinkage 0.05,
    step_count_smoothing 10, decay_rate 0.75, shrinkage_target = log(10 * eps0)."""
    step_size: float
    log_shrinkage_target: float
    error_sum: float = 0.0
    log_averaging_step: float = 0.0
    step: int = 0

    @classmethod
    def create(cls, eps0):
        return cls(step_size=float(eps0), log_shrinkage_target=math.log(10.0 * eps0))


def dual_averaging_update(st: DualAveragingState, accept_prob: float, num_adaptation_steps: int,
                          target=0.75, shrinkage=0.05, smoothing=10.0, decay=0.75):
    if st.step < num_adaptation_steps:
        st.error_sum += target - accept_prob
        t = st.step + 1.0
        log_x = st.log_shrinkage_target - math.sqrt(t) * st.error_sum / (shrinkage * (t + smoothing))
        eta = t ** (-decay)
        st.log_averaging_step = eta * log_x + (1.0 - eta) * st.log_averaging_step
        st.step_size = math.exp(log_x)
        if st.step + 1 == num_adaptation_steps:
            st.step_size = math.exp(st.log_averaging_step)
    st.step += 1
    return st


def hmc_chain(c: PosteriorConstants, model: str, z0, n_iter, n_leapfrog, eps0, seed, chain_id,
              num_adaptation_steps=0, min_temp=0.1, step0=0, grad="analytic", fixed_beta_temp=None):
    """Fixed-length HMC with the reference's sampler glue: per-iteration temperature
    beta_temp = schedule(step) (magi_v2.py:855-856), one scalar step size for all state parts
    (:364), identity mass, dual averaging over the first ``num_adaptation_steps`` iterations
    (:365-366), Metropolis accept on H = -lp + p.p/2.  The log-posterior at the current point is
    re-evaluated at the new temperature at the start of each iteration.  Randomness comes from the
    counter-based generator above so that the CUDA sampler can be checked draw-for-draw."""
    n, D = c.n, c.D
    P = len(z0) - n * D - D
    da = DualAveragingState.create(eps0)
    z = np.array(z0, dtype=np.float64, copy=True)
    out_z, out_acc, out_eps, out_lp = [], [], [], []
    for it in range(n_iter):
        bt = float(fixed_beta_temp) if fixed_beta_temp is not None else \
            float(logarithmic_temperature_schedule(step0 + it, min_temp))

        def vg(zz):
            X, s, tau = unpack_state(zz, n, D, P)
            if grad == "analytic":
                lp, gX, gs, gt = log_posterior_and_grad_analytic(X, s, tau, bt, c, model)
            else:
                lp, gX, gs, gt = log_posterior_and_grad_autograd(X, s, tau, bt, c)
            return lp, pack_state(gX, gs, gt)

        eps = da.step_size
        p0 = rng_normals(seed, chain_id, it, len(z))
        lp0, _ = vg(z)
        z1, p1, lp1, _, _ = leapfrog(z, p0, eps, n_leapfrog, vg)
        h0 = -lp0 + 0.5 * np.dot(p0, p0)
        h1 = -lp1 + 0.5 * np.dot(p1, p1)
        dH = h1 - h0
        acc_prob = 0.0 if not np.isfinite(dH) else min(1.0, math.exp(min(0.0, -dH)))
        u = rng_uniform(seed, chain_id, it)
        accepted = u < acc_prob
        if accepted:
            z = z1
        da = dual_averaging_update(da, acc_prob, num_adaptation_steps)
        out_z.append(z.copy()); out_acc.append(acc_prob); out_eps.append(eps)
        out_lp.append(lp1 if accepted else lp0)
    return np.array(out_z), np.array(out_acc), np.array(out_eps), np.array(out_lp)


# ---- NUTS -------------------------------------------------------
# tfp.mcmc.NoUTurnSampler (tensorflow-probability==0.24.0, mcmc/nuts.py; source not under
# /root/reference -- call sites magi_v2.py:360-366, :866-869) restated from the published algorithm
# (Hoffman & Gelman 2014 with Betancourt's multinomial sampling and the generalised U-turn
# criterion, which is what TFP implements): tree doubling up to max_tree_depth = 10, leaves weighted
# by exp(H0 - H), uniform progressive sampling inside a new subtree, biased progressive sampling
# between the old tree and the new subtree, U-turn checked on every complete dyadic sub-tree of the
# new subtree and on the merged tree with rho = sum of momenta, divergence when H - H0 > 1000.
# This version is RECURSIVE (one chain); the product (magi_v2_b200/nuts.py) is the iterative batched
# form with checkpoint memory -- the two are checked against each other draw for draw.

RNG_PURPOSE_NUTS_DEPTH = 2     # per doubling: word pair 0 -> direction, pair 1 -> subtree acceptance
RNG_PURPOSE_NUTS_LEAF = 3      # per leaf: multinomial selection inside the new subtree


def rng_uniform_pair(seed: int, chain_id: int, iteration: int, purpose: int, index: int):
    ctr = np.array([[index, chain_id, iteration, purpose]], dtype=np.uint32)
    key = np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    return float(_u53(r[:, 0], r[:, 1])[0]), float(_u53(r[:, 2], r[:, 3])[0])


def _logaddexp(a, b):
    return float(np.logaddexp(a, b))


def _no_uturn(rho, p_first, p_last):
    return (np.dot(rho, p_first) > 0.0) and (np.dot(rho, p_last) > 0.0)


def nuts_transition(z, eps, value_and_grad, seed, chain_id, iteration, max_tree_depth=10,
                    max_energy_diff=1000.0):
    """One NUTS transition from z.  Returns (z_new, lp_new, accept_stat, n_leapfrog, depth, diverged)."""
    S = len(z)
    lp0, g0 = value_and_grad(z)
    p0 = rng_normals(seed, chain_id, iteration, S)
    H0 = -lp0 + 0.5 * np.dot(p0, p0)
    stats = {"sum_acc": 0.0, "n_leaf": 0, "leaf_index": 0, "diverged": False}

    def build(zc, pc, gc, direction, depth):
        """2^depth leaves continuing from (zc, pc, gc).  Returns dict with the last state, rho, first / last
        momentum, log weight, proposal and `ok` (False = turning or diverged: nothing after it is built)."""
        if depth == 0:
            e = direction * eps
            ph = pc + 0.5 * e * gc
            zn = zc + e * ph
            lpn, gn = value_and_grad(zn)
            pn = ph + 0.5 * e * gn
            H = -lpn + 0.5 * np.dot(pn, pn)
            dE = H - H0
            if not np.isfinite(dE):
                dE = np.inf
            stats["sum_acc"] += min(1.0, math.exp(min(0.0, -dE)))
            stats["n_leaf"] += 1
            u_leaf, _ = rng_uniform_pair(seed, chain_id, iteration, RNG_PURPOSE_NUTS_LEAF, stats["leaf_index"])
            stats["leaf_index"] += 1
            div = dE > max_energy_diff
            stats["diverged"] |= bool(div)
            return {"z": zn, "p": pn, "g": gn, "rho": pn.copy(), "p_first": pn, "p_last": pn, "logw": -dE,
                    "prop": (zn, lpn), "ok": not div, "u": [u_leaf]}
        a = build(zc, pc, gc, direction, depth - 1)
        if not a["ok"]:
            return a
        b = build(a["z"], a["p"], a["g"], direction, depth - 1)
        logw = _logaddexp(a["logw"], b["logw"])
        out = dict(b)
        out["rho"] = a["rho"] + b["rho"]
        out["p_first"] = a["p_first"]
        out["logw"] = logw
        out["u"] = a["u"] + b["u"]
        out["prop"] = a["prop"]          # resolved below (sequential selection, leaf by leaf)
        out["ok"] = b["ok"] and _no_uturn(out["rho"], out["p_first"], out["p_last"])
        out["halves"] = (a, b)
        return out

    def leaves(t):
        if "halves" not in t:
            return [t]
        a, b = t["halves"]
        return leaves(a) + leaves(b)

    z_l = z_r = np.array(z, dtype=np.float64)
    p_l = p_r = p0
    g_l = g_r = g0
    rho = p0.copy()
    logw = 0.0
    prop_z, prop_lp = np.array(z, dtype=np.float64), lp0
    depth = 0
    while depth < max_tree_depth:
        u_dir, u_acc = rng_uniform_pair(seed, chain_id, iteration, RNG_PURPOSE_NUTS_DEPTH, depth)
        direction = 1.0 if u_dir < 0.5 else -1.0
        if direction > 0:
            t = build(z_r, p_r, g_r, direction, depth)
        else:
            t = build(z_l, p_l, g_l, direction, depth)
        depth += 1
        if not t["ok"]:
            break
        # uniform progressive selection inside the subtree, leaf by leaf: leaf i replaces the running
        # proposal with probability w_i / (w_0 + ... + w_i)
        run_w, sub_prop = -np.inf, None
        for lf in leaves(t):
            run_w = _logaddexp(run_w, lf["logw"])
            if math.log(lf["u"][0]) < lf["logw"] - run_w:
                sub_prop = lf["prop"]
        if math.log(u_acc) < t["logw"] - logw:
            prop_z, prop_lp = sub_prop
        logw = _logaddexp(logw, t["logw"])
        rho = rho + t["rho"]
        if direction > 0:
            z_r, p_r, g_r = t["z"], t["p"], t["g"]
        else:
            z_l, p_l, g_l = t["z"], t["p"], t["g"]
        if not _no_uturn(rho, p_l, p_r):
            break
    acc = stats["sum_acc"] / max(stats["n_leaf"], 1)
    return prop_z, prop_lp, acc, stats["n_leaf"], depth, stats["diverged"]


def nuts_chain(c: PosteriorConstants, model: str, z0, n_iter, eps0, seed, chain_id, num_adaptation_steps=0,
               min_temp=0.1, step0=0, fixed_beta_temp=None, max_tree_depth=10):
    """The reference's sampler stack (magi_v2.py:357-371, :852-879): NUTS(step 0.1) inside dual averaging
    (target 0.75, 0.8 * burn-in adaptation steps) inside the log-annealing wrapper."""
    n, D = c.n, c.D
    P = len(z0) - n * D - D
    da = DualAveragingState.create(eps0)
    z = np.array(z0, dtype=np.float64, copy=True)
    out_z, out_acc, out_eps, out_nleap = [], [], [], []
    for it in range(n_iter):
        bt = float(fixed_beta_temp) if fixed_beta_temp is not None else \
            float(logarithmic_temperature_schedule(step0 + it, min_temp))

        def vg(zz):
            X, s, tau = unpack_state(zz, n, D, P)
            lp, gX, gs, gt = log_posterior_and_grad_analytic(X, s, tau, bt, c, model)
            return lp, pack_state(gX, gs, gt)

        eps = da.step_size
        z, _, acc, nl, _, _ = nuts_transition(z, eps, vg, seed, chain_id, step0 + it, max_tree_depth)
        da = dual_averaging_update(da, acc, num_adaptation_steps)
        out_z.append(z.copy()); out_acc.append(acc); out_eps.append(eps); out_nleap.append(nl)
    return np.array(out_z), np.array(out_acc), np.array(out_eps), np.array(out_nleap)
