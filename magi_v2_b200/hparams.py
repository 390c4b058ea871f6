"""GP hyper-parameter initial fit, batched on the device.

Restates ``MAGI_v2._fit_kernel_hparams`` (magi_v2.py:538-691): Fourier-informed prior on phi2
(:549-565), then 1000 Adam steps (lr 0.01, :654) on softplus-transformed (phi1, sigma^2, phi2) (:631-642)
maximising, per component,
    log N(x_d ; mu_d 1, phi1 R(phi2) + (sigma^2 + jitter) I)                       (:594-597, jitter 1e-6)
  + log TN(phi1; 1e-4, 1000 sqrt(D)) + log TN(sigma^2; (0.1 sd)^2, 1000 sqrt(D)) + log TN(phi2; mu_phi2, sd_phi2 sqrt(D))
(:610-628; the reference's (D,1)+(D,) broadcast multiplies the whole objective by D, which Adam's
normalisation removes).  TFP autodiff is replaced by the closed-form gradient
    d ll / d h = 1/2 tr( (a a^T - S^-1) dS/dh ),   a = S^-1 (x - mu),
with dS/dphi1 = C/phi1, dS/dsigma^2 = I and dS/dphi2 = -C'_{ij} (s_i - s_j)/phi2, where C and C' come
from the library's Matern kernel (torch.ops.magi_b200.cov_build) and S^-1, log det S from the library's batched
Cholesky (`magi_b200_spd_inverse`, csrc/factor.cu); the remaining algebra is batched torch tensor code."""
from __future__ import annotations

import numpy as np

JITTER = 1e-6
NU = 2.01


def fourier_prior(X_filled: np.ndarray):
    """magi_v2.py:549-565.  X_filled [..., n, D] -> (mu_phi2 [..., D], sd_phi2 [..., D])."""
    z = np.fft.fft(X_filled, axis=-2)
    zmod = np.abs(z)
    n = X_filled.shape[-2]
    eff = zmod[..., 1:(n - 1) // 2 + 1, :] ** 2
    idxs = np.linspace(1, eff.shape[-2], eff.shape[-2]).reshape((-1, 1))
    freq = np.sum(idxs * eff, axis=-2) / np.sum(eff, axis=-2)
    mu_phi2 = 0.5 / freq
    return mu_phi2, (1 - mu_phi2) / 3


def objective_and_grad(v, grid, dt, xc, loc, scale, uniform_grid: bool = False):
    """Per (dataset, component): log N(x; mu 1, phi1 R(phi2) + (sigma^2 + jitter) I) + the three normal prior
    terms, as a function of the softplus pre-activations v [3,B,D] = (phi1, phi2, sigma^2), and the gradient of
    its NEGATIVE (the Adam loss) with respect to v.  Closed form:
    d ll / d h = 1/2 tr((a a^T - S^-1) dS/dh), a = S^-1 (x - mu); dS/dphi1 = C/phi1, dS/dsigma^2 = I,
    dS/dphi2 = -C'_{ij} (s_i - s_j)/phi2 with C, C' from torch.ops.magi_b200.cov_build."""
    import torch
    from . import ops
    n = grid.shape[0]
    eye = torch.eye(n, dtype=torch.float64, device=v.device)
    h = torch.nn.functional.softplus(v)
    phi1, phi2, sig2 = h[0].contiguous(), h[1].contiguous(), h[2]
    C, Cp, _ = ops.cov_build(grid, phi1, phi2, NU, uniform_grid)
    S = C + (sig2 + JITTER)[..., None, None] * eye
    Sinv, logdet, info = ops.spd_inverse(S.contiguous())     # the library's blocked Cholesky (csrc/factor.cu)
    if int(info.abs().max()) != 0:
        bad = torch.nonzero(info)[:4].tolist()
        raise np.linalg.LinAlgError(f"GP covariance not positive definite for (dataset, component) {bad}")
    a = (Sinv @ xc[..., None])[..., 0]                                          # [B,D,n]
    W = a[..., :, None] * a[..., None, :] - Sinv
    g_phi1 = 0.5 * (W * C).sum(dim=(-1, -2)) / phi1
    g_sig2 = 0.5 * torch.diagonal(W, dim1=-2, dim2=-1).sum(-1)
    g_phi2 = 0.5 * (W * (-Cp * dt)).sum(dim=(-1, -2)) / phi2
    g = torch.stack([g_phi1, g_phi2, g_sig2])
    g = g - (h - loc) / scale ** 2                                              # truncated-normal priors
    g = -g * torch.sigmoid(v)                                                   # loss = -log_prob; softplus chain rule
    ll = -0.5 * (xc * a).sum(-1) - 0.5 * logdet - 0.5 * n * np.log(2 * np.pi)
    obj = ll - 0.5 * (((h - loc) / scale) ** 2).sum(0)
    return obj, g


def fit_kernel_hparams(I: np.ndarray, X_filled: np.ndarray, device="cuda:0", num_iters: int = 1000,
                       lr: float = 0.01, verbose: bool = False):
    """I [n]; X_filled [B, n, D] (no NaNs).  Returns dict of phi1s, phi2s, sigma_sqs, each [B, D]."""
    import torch
    from . import ops

    if not torch.cuda.is_available():
        raise RuntimeError("magi_v2_b200 needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device)
    B, n, D = X_filled.shape
    mu_phi2, sd_phi2 = fourier_prior(X_filled)                                  # [B,D]
    sd = X_filled.std(axis=1)                                                   # [B,D]
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    inv_softplus = lambda a: np.log(np.expm1(a))
    # trainable pre-activations, order (phi1, phi2, sigma^2) as :645
    v = T(np.stack([inv_softplus(sd ** 2), inv_softplus(mu_phi2), inv_softplus((0.1 * sd) ** 2)]))   # [3,B,D]
    loc = T(np.stack([np.full((B, D), 1e-4), mu_phi2, (0.1 * sd) ** 2]))
    scale = T(np.stack([np.full((B, D), 1000.0 * np.sqrt(D)), sd_phi2 * np.sqrt(D),
                        np.full((B, D), 1000.0 * np.sqrt(D))]))
    grid = T(I)
    steps = np.diff(np.asarray(I, dtype=np.float64))
    uniform = bool(np.allclose(steps, steps[0], rtol=1e-10, atol=0.0))          # Toeplitz kernel of cov_build
    dt = grid[:, None] - grid[None, :]                                         # s_i - s_j
    x = T(np.transpose(X_filled, (0, 2, 1)))                                    # [B,D,n]
    xc = x - x.mean(dim=-1, keepdim=True)                                       # mean_fn = column mean (:559, :589)
    m1 = torch.zeros_like(v)
    m2 = torch.zeros_like(v)
    b1, b2, eps = 0.9, 0.999, 1e-7                                              # tf_keras Adam defaults
    for t in range(1, num_iters + 1):
        obj, g = objective_and_grad(v, grid, dt, xc, loc, scale, uniform)
        g = D * g          # the reference's loss is the [D, D] broadcast of :603-607 summed by tape.gradient: D x the
                           # per-component objective (Adam is scale-free up to its epsilon; kept for step-for-step parity)
        m1 = b1 * m1 + (1 - b1) * g
        m2 = b2 * m2 + (1 - b2) * g * g
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        v = v - lr_t * m1 / (torch.sqrt(m2) + eps)
        if verbose and (t % 100 == 0 or t == 1):
            print(f"[hparams] iter {t}: mean objective {float(obj.mean()):.4f}")
    h = torch.nn.functional.softplus(v).cpu().numpy()
    return {"phi1s": h[0], "phi2s": h[1], "sigma_sqs": h[2]}
