"""Initialisation of completely unobserved components (magi_v2.py:182-250).

The reference fits (X_unobs, thetas) jointly by gradient matching before the GP hyper-parameters of the unobserved
components can be fitted: the observed components are fixed at their smoothed interpolants, and Adam (tf_keras
defaults, lr 0.01, 10 000 steps from thetas = 1 and X_unobs ~ N(mean, sd) of the interpolated data) minimises
    sum_{i=1..n-2, d} ( f_d(X_i, thetas) - (X_{i+1,d} - X_{i-1,d}) / (2 dx) )^2        (:200-217)
This is set-up (a few thousand scalars, once per fit), not the sampling path: torch FP64 autograd on the device the
model lives on, with the registry's right-hand side evaluated on tensors."""
from __future__ import annotations

import numpy as np


def fit_unobserved(model, I: np.ndarray, X_smoothed_obs: np.ndarray, observed_components, unobserved_components,
                   X_interp_obs: np.ndarray, num_iters: int = 10000, lr: float = 0.01, seed=None, device="cpu"):
    """Returns (X_unobs [n, D_unobs], thetas [P], loss_first, loss_last)."""
    import torch
    dev = torch.device(device)
    n = X_smoothed_obs.shape[0]
    obs, unobs = list(observed_components), list(unobserved_components)
    order = np.argsort(np.concatenate([obs, unobs]))                                   # proper_order, :50
    rng = np.random.default_rng(seed)
    mu0 = X_interp_obs.mean()                                                          # :220
    sd0 = float((X_interp_obs.std(axis=0) ** 2).mean() ** 0.5)                         # :221
    Xu = torch.tensor(rng.normal(mu0, sd0, (n, len(unobs))), dtype=torch.float64, device=dev, requires_grad=True)  # :224-227
    th = torch.ones(model.P, dtype=torch.float64, device=dev, requires_grad=True)      # :228
    Xo = torch.as_tensor(np.ascontiguousarray(X_smoothed_obs), dtype=torch.float64, device=dev)
    It = torch.as_tensor(np.asarray(I, dtype=np.float64).reshape(-1, 1), device=dev)
    idx = torch.as_tensor(order, device=dev)
    dx2 = 2.0 * float(I.reshape(-1)[1] - I.reshape(-1)[0])                             # equally spaced I assumed, :213
    params = [Xu, th]
    m1 = [torch.zeros_like(p) for p in params]
    v1 = [torch.zeros_like(p) for p in params]
    b1, b2, eps = 0.9, 0.999, 1e-7                                                     # tf_keras Adam defaults
    first = None
    for t in range(1, num_iters + 1):
        X_full = torch.cat([Xo, Xu], dim=1)[:, idx]                                    # :202-203
        f_vals = model.f_vec(It, X_full, th)                                           # :206
        f_diff = (X_full[2:] - X_full[:-2]) / dx2                                      # :213
        loss = ((f_vals[1:-1] - f_diff) ** 2).sum()                                    # :216
        grads = torch.autograd.grad(loss, params)
        lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, m1, v1):
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                p.sub_(lr_t * m / (v.sqrt() + eps))
        if t == 1:
            first = float(loss.detach())
    last = float(loss.detach())
    return Xu.detach().cpu().numpy().copy(), th.detach().cpu().numpy().copy(), first, last
