"""Shared test helpers: build oracle constants for small problems and mirror them on the device."""
import os

import numpy as np

from oracle import magi_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def synth_constants(model_name, seed, N=9, disc=1, band=6, T=2.0, nan_frac=0.2):
    """A small seeded posterior (oracle constants) for a registry model."""
    model = mo.MODELS[model_name]
    rng = np.random.default_rng(seed)
    D = model.D
    ts = np.linspace(0, T, N)
    X_obs = np.abs(rng.normal(0.3, 0.2, size=(N, D)))
    X_obs[rng.uniform(size=X_obs.shape) < nan_frac] = np.nan
    X_obs[0] = 0.2
    X_obs[-1] = 0.4
    phi1 = rng.uniform(0.005, 0.05, D)
    phi2 = rng.uniform(0.2, 0.6, D)
    c = mo.make_constants(ts, X_obs, disc, phi1, phi2, band, model.f_vec)
    return c


def seir_vignette_constants(which=0, model_name="seir3", band=80):
    """vignette.ipynb settings on one of the 21 thinned SEIR datasets (tests/golden/seir_datasets.npz),
    with illustrative kernel hyper-parameters."""
    g = load_golden("seir_datasets.npz")
    ts = g["ts_obs"]
    X4 = g["X_obs"][which].copy()
    X4[X4 < 0.0] = 0.0                                  # vignette.ipynb:112-113
    model = mo.MODELS[model_name]
    if model_name == "seir3":
        X_obs = X4[:, 1:]
        phi1, phi2 = (0.0085, 0.034, 0.024), (0.375, 0.23, 0.109)
    else:
        X_obs = X4
        phi1, phi2 = (0.05, 0.0085, 0.034, 0.024), (0.3, 0.375, 0.23, 0.109)
    return mo.make_constants(ts, X_obs, 1, phi1, phi2, band, model.f_vec), ts, X_obs


def random_state(c, model_name, rng, R, jitter=0.02):
    model = mo.MODELS[model_name]
    n, D, P = c.n, model.D, model.P
    y, mask = c.dense_y_mask()
    base = np.where(mask > 0, y, 0.0)
    # a plausible trajectory: interpolate the observations
    Xd = np.where(mask > 0, y, np.nan)
    X0 = mo.linear_interpolate(Xd)
    X = X0[None] + jitter * rng.standard_normal((R, n, D))
    s = rng.normal(-4, 1, (R, D))
    tau = rng.normal(0.5, 1.0, (R, P))
    return X, s, tau


def matrix_bandwidth(consts):
    """Largest |i-j| holding a non-zero in any of C^-1, m, K^-1 of the given constants (the bandsize the
    reference's band_part left, magi_v2.py:271-274), or None when the matrices are dense."""
    bw = 0
    n = consts[0].n
    for c in consts:
        for A in (c.C_d_invs, c.m_ds, c.K_d_invs):
            i, j = np.nonzero(np.abs(A).sum(axis=0))
            bw = max(bw, int(np.abs(i - j).max()))
    return None if bw >= n - 1 else bw


def device_problem(consts, model_name, device, band="auto"):
    """List of oracle PosteriorConstants (same n, model) -> magi_v2_b200.ops.PosteriorProblem.
    band="auto": tell the kernels the true bandwidth of the matrices; None: treat them as dense."""
    import torch
    from magi_v2_b200 import ops
    if band == "auto":
        band = matrix_bandwidth(consts)

    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    Cinv = T(np.stack([c.C_d_invs for c in consts]))
    m = T(np.stack([c.m_ds for c in consts]))
    Kinv = T(np.stack([c.K_d_invs for c in consts]))
    packed = ops.pack_matrices(Cinv, m, Kinv)
    ys, masks = zip(*[c.dense_y_mask() for c in consts])
    prob = ops.PosteriorProblem(
        model_name, packed,
        mu=T(np.stack([c.mu_ds for c in consts])),
        y=T(np.stack(ys)), mask=T(np.stack(masks), torch.uint8),
        N_ds=T(np.stack([c.N_ds.astype(np.float64) for c in consts])),
        beta=T(np.array([c.beta for c in consts])),
        LB=T(np.stack([c.sigma_sqs_LB for c in consts])), n=consts[0].n, band=band)
    return prob


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def chol_solve_longdouble(A, Bm=None):
    """Extended-precision (x87 80-bit) Cholesky inverse / solve in plain numpy: the 'truth' against
    which both the reference's SVD route and the CUDA Cholesky route are measured."""
    A = np.asarray(A, dtype=np.longdouble)
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        d = A[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(d)
        if j + 1 < n:
            L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    rhs = np.eye(n, dtype=np.longdouble) if Bm is None else np.asarray(Bm, dtype=np.longdouble)
    Y = np.zeros_like(rhs)
    for i in range(n):
        Y[i] = (rhs[i] - L[i, :i] @ Y[:i]) / L[i, i]
    Z = np.zeros_like(rhs)
    for i in range(n - 1, -1, -1):
        Z[i] = (Y[i] - L[i + 1:, i] @ Z[i + 1:]) / L[i, i]
    return Z


def matern_truth_longdouble(I, phi1, phi2, nu):
    """(C^-1, m, K, K^-1) in extended precision from double-precision Matern blocks."""
    from oracle import magi_oracle as mo
    Kap, pK, Kpp = mo.matern_blocks(I, phi1, phi2, nu)
    # the reference's Kappa_pp is the noisy one (see matern_blocks_roundoff_scale): rebuild it as the
    # exact Toeplitz-consistent function of the lag from the first row where s = 0 has no cancellation
    Cinv = chol_solve_longdouble(Kap)
    pK_l = np.asarray(pK, dtype=np.longdouble)
    m = pK_l @ Cinv
    K = np.asarray(Kpp, dtype=np.longdouble) - m @ pK_l.T
    K = (K + K.T) / 2
    Kinv = chol_solve_longdouble(K)
    return Cinv, m, K, Kinv
