"""Generate the committed fixtures under tests/golden/ (run in the build container only:
needs /root/reference).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

    python -m oracle.make_golden

Fixtures:
  build_kat.npz       GENUINE reference `_build_matrices` (magi_v2.py:774-823) outputs C, m, K for small
                      grids, plus row/column probes and norms at n = 161 / 321; Matern blocks
                      (Kappa, p_Kappa, Kappa_pp) from the restatement (bit-identical C pins it).
  grid_kat.npz        GENUINE `_discretize` / `_linear_interpolate` outputs.
  seir_datasets.npz   the 21 SEIR CSVs of the reference, thinned as vignette.ipynb:100-113 does
                      (t <= 4, every 50th row): ts_obs [81], X_obs [21, 81, 4] (S,E,I,R _obs columns),
                      X_true [21, 81, 4], names.  Data, not code.
  logpost_kat.npz     seeded inputs and the restated log-posterior value + autograd gradient
                      (oracle.magi_oracle) for every registry model at small n; matrices come from
                      the genuine `_build_matrices` + tf_pinv stand-in + band.
  init_kat.npz        initial-fit stages on the vignette data (SEIR seed 0, E/I/R, discretization 1): GENUINE
                      `cv_cubic_smoother` output (magi_v2.py:695-770); from the restatement
                      (oracle.init_oracle; TFP / tf_keras not installable, "parity unpinned"): Fourier prior,
                      the hyper-parameter objective and its gradient at the start, 25 Adam steps of the fit,
                      thetas_init in both layouts (1500 Adam steps, illustrative phi).
  vignette_fit.npz    `initial_fit` of the vignette through the restatement: fitted (phi1, phi2, sigma^2) after the
                      reference's 1000 Adam steps, thetas_init after its 10 000 (both layouts), smoothed Xhat_init.
  vignette_chains.npz written by `python -m oracle.vignette_study --golden`: per-chain theta / sigma^2 means of the
                      reference's sampler stack (1000 + 1000 NUTS transitions, annealing schedule) on the C oracle.
"""
import glob
import os

import numpy as np
import pandas as pd

from . import magi_oracle as mo
from .ref_loader import REFERENCE_DIR, reference_object

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_kat(ref):
    out = {}
    cases = [("appB", np.array([0, .025, .05, .075, .1]), 0.02, 0.23, 2.01),
             ("n21", np.linspace(0, 1, 21), 0.0085, 0.375, 2.01),
             ("n41", np.linspace(0, 4, 41), 0.034, 0.23, 2.01),
             ("n33nu25", np.linspace(0, 2, 33), 0.024, 0.109, 2.5),
             ("n17ragged", np.sort(np.random.default_rng(1).uniform(0, 2, 17)), 0.02, 0.3, 2.01)]
    for name, I, p1, p2, v in cases:
        C, m, K = ref._build_matrices(I.reshape(-1, 1), p1, p2, v)
        Kap, pK, Kpp = mo.matern_blocks(I, p1, p2, v)
        assert np.array_equal(Kap, C)
        out[f"{name}_I"] = I; out[f"{name}_hp"] = np.array([p1, p2, v])
        out[f"{name}_C"] = C; out[f"{name}_m"] = m; out[f"{name}_K"] = K
        out[f"{name}_pK"] = pK; out[f"{name}_Kpp"] = Kpp
    for n, p2 in ((161, 0.375), (161, 0.109), (321, 0.23)):
        I = np.linspace(0, 4, n)
        C, m, K = ref._build_matrices(I.reshape(-1, 1), 0.0085, p2, 2.01)
        Kap, pK, Kpp = mo.matern_blocks(I, 0.0085, p2, 2.01)
        tag = f"n{n}_phi2_{p2}"
        out[f"{tag}_hp"] = np.array([0.0085, p2, 2.01])
        for nm, A in (("C", C), ("m", m), ("K", K), ("pK", pK), ("Kpp", Kpp)):
            out[f"{tag}_{nm}_row0"] = A[0]; out[f"{tag}_{nm}_rowmid"] = A[n // 2]
            out[f"{tag}_{nm}_diag"] = np.diag(A); out[f"{tag}_{nm}_fro"] = np.linalg.norm(A)
    np.savez_compressed(os.path.join(OUT, "build_kat.npz"), **out)


def grid_kat(ref):
    rng = np.random.default_rng(2)
    ts = np.linspace(0, 4, 11)
    X = rng.normal(size=(11, 3))
    X[rng.uniform(size=X.shape) < 0.3] = np.nan
    X[0] = 1.0; X[-1] = 2.0
    out = {"ts": ts, "X": X}
    for disc in (0, 1, 2):
        I, Xd = ref._discretize(ts, X, disc)
        out[f"I_d{disc}"] = I; out[f"Xd_d{disc}"] = Xd
        out[f"Xi_d{disc}"] = ref._linear_interpolate(Xd)
    np.savez_compressed(os.path.join(OUT, "grid_kat.npz"), **out)


def seir_datasets():
    files = [os.path.join(REFERENCE_DIR, "data", "SEIR_seed=0.csv")]
    for a in ("0.05", "0.15"):
        for s in range(10):
            files.append(os.path.join(REFERENCE_DIR, "data", f"SEIR_beta=6_gamma=0.6_sigma=1.8_alpha={a}_seed={s}.csv"))
    Xo, Xt, names, ts = [], [], [], None
    for f in files:
        raw = pd.read_csv(f).query("t <= 4.0")                                     # vignette.ipynb:104
        obs = raw.iloc[::int((raw.index.shape[0] - 1) / (20 * 4.0))]               # :105
        ts = obs.t.values.astype(np.float64)
        Xo.append(obs[["S_obs", "E_obs", "I_obs", "R_obs"]].to_numpy().astype(np.float64))
        Xt.append(obs[["S_true", "E_true", "I_true", "R_true"]].to_numpy().astype(np.float64))
        names.append(os.path.basename(f))
    np.savez_compressed(os.path.join(OUT, "seir_datasets.npz"), ts_obs=ts, X_obs=np.array(Xo), X_true=np.array(Xt),
                        names=np.array(names))


def logpost_kat(ref):
    out = {}
    rng = np.random.default_rng(3)
    for name, model in mo.MODELS.items():
        D, P = model.D, model.P
        N = 9
        ts = np.linspace(0, 2, N)
        X_obs = np.abs(rng.normal(0.3, 0.2, size=(N, D)))
        X_obs[rng.uniform(size=X_obs.shape) < 0.2] = np.nan
        X_obs[0] = 0.2; X_obs[-1] = 0.4
        phi1 = rng.uniform(0.005, 0.05, D); phi2 = rng.uniform(0.2, 0.6, D)
        I, _ = mo.discretize(ts, X_obs, 1)
        n = I.shape[0]
        band = 6
        mats = [np.zeros((D, n, n)) for _ in range(3)]
        for d in range(D):
            C_d, m_d, K_d = ref._build_matrices(I, phi1[d], phi2[d], 2.01)       # genuine reference
            mats[0][d], mats[1][d], mats[2][d] = mo.tf_pinv(C_d), m_d, mo.tf_pinv(K_d)
        mats = [mo.band_part(A, band) for A in mats]
        c = mo.make_constants(ts, X_obs, 1, phi1, phi2, band, model.f_vec, matrices=mats)
        X = mo.linear_interpolate(mo.discretize(ts, X_obs, 1)[1]) + 0.02 * rng.standard_normal((n, D))
        s = rng.normal(-4, 1, D); tau = rng.normal(0.5, 1.0, P); bt = 0.37
        lp, gX, gs, gt = mo.log_posterior_and_grad_autograd(X, s, tau, bt, c)
        lp_np = mo.log_posterior(X, s, tau, bt, c)
        assert abs(lp - lp_np) <= 1e-12 * abs(lp)
        for k, v in dict(ts=ts, X_obs=X_obs, phi1=phi1, phi2=phi2, band=np.array(band), Cinv=mats[0], m=mats[1],
                         Kinv=mats[2], X=X, s=s, tau=tau, bt=np.array(bt), lp=np.array(lp), gX=gX, gs=gs,
                         gt=gt).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "logpost_kat.npz"), **out)


def init_kat(ref):
    import torch
    from . import init_oracle as io
    g = np.load(os.path.join(OUT, "seir_datasets.npz"))
    ts, X = g["ts_obs"], g["X_obs"][0][:, 1:].copy()
    X[X < 0.0] = 0.0
    I, Xd = mo.discretize(ts, X, 1)
    Xi = mo.linear_interpolate(Xd)
    out = {"X_interp": Xi, "I": I}
    out["smoothed_genuine"] = ref.cv_cubic_smoother(I, Xi)                        # genuine reference code
    assert np.allclose(out["smoothed_genuine"], io.cv_cubic_smoother(I, Xi), rtol=0, atol=1e-13)
    mu2, sd2 = io.fourier_prior(Xi)
    out["mu_phi2"], out["sd_phi2"] = mu2, sd2
    obj = io.HparamObjective(I, Xi)
    v = obj.initial_variables()
    loss, grads = obj.loss_and_grads(v)
    out["hp_v0"] = np.stack([a.detach().numpy() for a in v])                      # (phi1, phi2, sigma^2) pre-activations
    out["hp_loss0"] = loss.numpy()                                                 # [D, D]
    out["hp_grad0"] = np.stack([a.numpy() for a in grads])
    trace = []
    io.fit_kernel_hparams(I, Xi, num_iters=25, trace=trace)
    out["hp_trace25"] = np.array([np.stack(t) for t in trace])                    # [25, 3, D]
    phi1, phi2 = np.array([0.0085, 0.034, 0.024]), np.array([0.375, 0.23, 0.109])
    dense = mo.kernel_matrices(I, phi1, phi2, None)
    out["ti_phi1"], out["ti_phi2"] = phi1, phi2
    for layout in ("reference", "transpose"):
        out[f"thetas_init_{layout}_1500"] = io.fit_thetas_init(I, Xi, Xi.mean(axis=0), dense[1], dense[2], mo.f_seir3, 3,
                                                               num_iters=1500, layout=layout)
    np.savez_compressed(os.path.join(OUT, "init_kat.npz"), **out)


def vignette_fit(hp=None):
    """The whole of `initial_fit` on the vignette data through the restatement (oracle.init_oracle): 1000 Adam steps of
    the hyper-parameter fit, 10 000 of the theta initialisation in both layouts, the smoothed start."""
    from . import init_oracle as io
    g = np.load(os.path.join(OUT, "seir_datasets.npz"))
    ts, X = g["ts_obs"], g["X_obs"][0][:, 1:].copy()
    X[X < 0.0] = 0.0
    I, Xd = mo.discretize(ts, X, 1)
    if hp is None:
        hp = io.fit_kernel_hparams(I, mo.linear_interpolate(Xd))
    out = {"phi1s": hp["phi1s"], "phi2s": hp["phi2s"], "sigma_sqs": hp["sigma_sqs"]}
    for layout in ("reference", "transpose"):
        fit = io.initial_fit(ts, X, 1, 80, mo.f_seir3, 3, hparams=hp, theta_layout=layout)
        out[f"thetas_init_{layout}"] = fit["thetas_init"]
    out["Xhat_init"] = fit["Xhat_init"]
    out["sigma_sqs_LB"] = fit["constants"].sigma_sqs_LB
    np.savez_compressed(os.path.join(OUT, "vignette_fit.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = reference_object()
    build_kat(ref)
    grid_kat(ref)
    seir_datasets()
    logpost_kat(ref)
    init_kat(ref)
    vignette_fit()
    for f in sorted(glob.glob(os.path.join(OUT, "*.npz"))):
        print(f, os.path.getsize(f))


if __name__ == "__main__":
    main()
