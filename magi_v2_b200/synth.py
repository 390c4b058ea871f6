"""Synthetic SEIR-shaped workloads and the batched front end
that turns raw observations of many datasets into device-resident posterior constants.

The front end restates, vectorised over datasets, the bookkeeping of ``MAGI_v2.__init__`` /
``initial_fit`` that defines the constants of the log-posterior (magi_v2.py:53, :85-114, :299-300);
the kernel matrices come from the library (cov_build -> factor_derive -> pack_matrices)."""
from __future__ import annotations

import numpy as np

from . import models as _models

SEIR_TRUTH = np.array([6.0, 0.6, 1.8])          # beta, gamma, sigma of the reference's data/*.csv
SEIR_X0 = np.array([0.99, 0.01, 0.0, 0.0])


def simulate(model: str, thetas: np.ndarray, x0: np.ndarray, t_max: float, dt: float = 1e-3):
    """RK4 trajectories for a batch of parameter vectors: thetas [B,P] -> (t [T+1], X [B,T+1,D])."""
    m = _models.REGISTRY[model]
    B = thetas.shape[0]
    steps = int(round(t_max / dt))
    X = np.empty((B, steps + 1, m.D))
    x = np.broadcast_to(np.asarray(x0, dtype=np.float64), (B, m.D)).copy()
    X[:, 0] = x
    th = thetas.T                                  # [P,B]: f_vec broadcasts th[k] over rows
    def rhs(z):                                    # z [B,D] -> [B,D], per-row parameters
        return m.f_vec(None, z, [th[k][:, None] for k in range(m.P)])

    for i in range(steps):
        k1 = rhs(x)
        k2 = rhs(x + 0.5 * dt * k1)
        k3 = rhs(x + 0.5 * dt * k2)
        k4 = rhs(x + dt * k3)
        x = x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        X[:, i + 1] = x
    return np.linspace(0.0, t_max, steps + 1), X


def seir_sweep(B: int, seed0: int = 0, model: str = "seir4", n_obs: int = 81, t_max: float = 4.0):
    """Config 4: per dataset theta log-uniform within +-30 % of (6, 0.6, 1.8), noise alpha in {0.05, 0.15}
    (sd = alpha * range of the true component, as in the reference's CSVs), dataset b seeded by seed0 + b.
    Returns dict(ts_obs [N], X_obs [B,N,D], thetas_true [B,P], alpha [B], X_true [B,N,D])."""
    m = _models.REGISTRY[model]
    thetas = np.empty((B, 3))
    alpha = np.empty(B)
    noise = np.empty((B, n_obs, 4))
    for b in range(B):
        rng = np.random.default_rng(seed0 + b)
        thetas[b] = SEIR_TRUTH * np.exp(rng.uniform(np.log(0.7), np.log(1.3), 3))
        alpha[b] = 0.05 if (seed0 + b) % 2 == 0 else 0.15
        noise[b] = rng.standard_normal((n_obs, 4))
    t, X = simulate("seir4", thetas, SEIR_X0, t_max)
    stride = (len(t) - 1) // (n_obs - 1)
    Xt = X[:, ::stride]
    rngs = Xt.max(axis=1, keepdims=True) - Xt.min(axis=1, keepdims=True)
    Xo = Xt + alpha[:, None, None] * rngs * noise
    Xo[Xo < 0.0] = 0.0                              # vignette.ipynb:112-113
    if model == "seir3":
        Xo, Xt = Xo[..., 1:], Xt[..., 1:]
    return {"ts_obs": t[::stride], "X_obs": Xo, "thetas_true": thetas, "alpha": alpha, "X_true": Xt}


def batch_constants(ts_obs: np.ndarray, X_obs: np.ndarray, discretization: int):
    """Vectorised magi_v2.py:53, :85-100, :105, :114 for fully observed datasets on a common time base:
    X_obs [B,N,D] -> dict(I [n], y [B,n,D], mask [B,n,D] u8, N_ds [B,D], beta [B], Xhat [B,n,D], mu [B,D])."""
    B, N, D = X_obs.shape
    stride = 2 ** discretization
    n = stride * (N - 1) + 1
    idx = np.arange(n)
    I = np.interp(idx, idx[::stride], np.asarray(ts_obs, dtype=np.float64))
    y = np.zeros((B, n, D))
    mask = np.zeros((B, n, D), dtype=np.uint8)
    obs = ~np.isnan(X_obs)
    y[:, ::stride] = np.where(obs, X_obs, 0.0)
    mask[:, ::stride] = obs
    N_ds = obs.sum(axis=1).astype(np.float64)
    beta = (D * n) / N_ds.sum(axis=1)
    Xhat = np.empty((B, n, D))
    for b in range(B):
        for d in range(D):
            have = mask[b, :, d] > 0
            Xhat[b, :, d] = np.interp(idx, idx[have], y[b, have, d])
    return {"I": I, "y": y, "mask": mask, "N_ds": N_ds, "beta": beta, "Xhat": Xhat, "mu": Xhat.mean(axis=1)}


def device_problem(model: str, I, phi1, phi2, y, mask, N_ds, beta, mu, LB, bandsize, device, nu: float = 2.01,
                   chunk: int = 512, uniform_grid: bool = True, keep_matrices: bool = False):
    """Kernel matrices for all datasets on the device (in chunks to bound the transient memory) and
    the PosteriorProblem that holds every constant of the log-posterior.  `keep_matrices` also keeps the
    dense UN-banded m and K^-1 ([B,D,n,n] each, `prob.kept_matrices`) for the theta initialisation."""
    import torch
    from . import ops
    dev = torch.device(device)
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    B, D = phi1.shape
    n = len(I)
    npad = (n + 7) // 8 * 8
    packed = torch.empty(B * D * 3 * npad * npad, dtype=torch.float64, device=dev)
    I_d, p1, p2 = T(I), T(phi1), T(phi2)
    band = -1 if bandsize is None else int(bandsize)
    infos, kept = [], []
    per = D * 3 * npad * npad
    keep_band = None
    if keep_matrices and band >= 0:
        # the theta initialisation needs the UN-banded m, K^-1 (magi_v2.py:132-179 runs before band_part, :271-274):
        # factorise dense, keep those, then zero outside the band before packing
        ii = torch.arange(n, device=dev)
        keep_band = (ii[:, None] - ii[None, :]).abs() <= band
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        C, Cp, Cpp = ops.cov_build(I_d, p1[b0:b1].contiguous(), p2[b0:b1].contiguous(), nu, uniform_grid)
        Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, -1 if keep_band is not None else band, 0.0)
        if keep_matrices:
            kept.append((m, Kinv))
        if keep_band is not None:
            zero = torch.zeros((), dtype=torch.float64, device=dev)
            Cinv, m, Kinv = (torch.where(keep_band, a, zero) for a in (Cinv, m, Kinv))
        packed[b0 * per:b1 * per] = ops.pack_matrices(Cinv, m, Kinv)
        infos.append(info)
        del C, Cp, Cpp, Cinv, m, Kinv
    info = torch.cat(infos)
    if LB is None:                                  # set later (PosteriorProblem.set_LB) once Xhat_init is smoothed
        LB = np.zeros((B, D))
    prob = ops.PosteriorProblem(model, packed, mu=T(mu), y=T(y), mask=T(mask, torch.uint8), N_ds=T(N_ds),
                                beta=T(beta), LB=T(LB), n=n, band=bandsize)
    if keep_matrices:
        prob.kept_matrices = (torch.cat([k[0] for k in kept]), torch.cat([k[1] for k in kept]))
    return prob, info


def sweep_problem(B: int, R: int, device, seed0: int = 0, model: str = "seir4", bandsize=80, chunk: int = 512):
    """Everything config 4 needs on one device: the problem constants and initial chain states (host)."""
    data = seir_sweep(B, seed0, model)
    c = batch_constants(data["ts_obs"], data["X_obs"], 1)
    D = c["mu"].shape[1]
    rng = np.random.default_rng(10_000_019 + seed0)
    phi1 = rng.uniform(0.005, 0.05, (B, D))
    phi2 = rng.uniform(0.1, 0.4, (B, D))
    sd = c["Xhat"].std(axis=1)
    LB = (0.01 * sd) ** 2                                                      # magi_v2.py:299-300
    prob, info = device_problem(model, c["I"], phi1, phi2, c["y"], c["mask"], c["N_ds"], c["beta"], c["mu"], LB,
                                bandsize, device, chunk=chunk)
    X0 = c["Xhat"][:, None] + 0.01 * sd[:, None, None, :] * rng.standard_normal((B, R, c["Xhat"].shape[1], D))
    rngs = data["X_true"].max(axis=1) - data["X_true"].min(axis=1)
    if rngs.shape[1] != D:
        rngs = rngs[:, -D:]
    sig2 = (data["alpha"][:, None] * rngs) ** 2
    s0 = np.where(sig2 > LB, np.log(np.expm1(np.maximum(sig2 - LB, 1e-300))), -5.0)
    th0 = data["thetas_true"] * np.exp(rng.uniform(-0.1, 0.1, data["thetas_true"].shape))
    tau0 = np.log(np.expm1(th0))
    state = {"X": X0, "sig_pre": np.repeat(s0[:, None], R, axis=1), "th_pre": np.repeat(tau0[:, None], R, axis=1)}
    data.update(phi1=phi1, phi2=phi2, LB=LB, consts=c, bandsize=bandsize)     # what a checker needs to rebuild a dataset
    return prob, info, state, data
