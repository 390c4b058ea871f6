"""Parity of the CUDA log-posterior + gradient (C-ABI magi_b200_logpost_grad) with the oracle's
op-for-op restatement of magi_v2.py:308-348 + autograd.  Tolerance: 1e-9 relative (BASELINE.json
north_star), FP64, given identical C^-1, m, K^-1."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests.helpers import device_problem, load_golden, random_state, relerr, seir_vignette_constants, synth_constants

pytestmark = pytest.mark.gpu
TOL = 1e-9


PATHS = ["cta", "wide"]      # magi_b200_logpost_grad / magi_b200_logpost_grad_wide: same function, two grid shapes


def _run(prob, X, s, tau, bt, device, path="cta"):
    import torch
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
    lp, gX, gs, gt = prob.logpost_grad(T(X), T(s), T(tau), T(bt), path=path)
    torch.cuda.synchronize()
    return lp.cpu().numpy(), gX.cpu().numpy(), gs.cpu().numpy(), gt.cpu().numpy()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("model", ["seir3", "seir4", "sirw", "lorenz96"])
@pytest.mark.parametrize("R", [1, 8, 11, 20])
def test_small_batches_match_oracle(model, R, path, cuda_device):
    B = 3
    rng = np.random.default_rng(10 + R)
    consts = [synth_constants(model, seed=100 + b) for b in range(B)]
    prob = device_problem(consts, model, cuda_device)
    st = [random_state(c, model, rng, R) for c in consts]
    X = np.stack([a[0] for a in st]); s = np.stack([a[1] for a in st]); tau = np.stack([a[2] for a in st])
    bt = rng.uniform(0.1, 1.5, (B, R))
    lp, gX, gs, gt = _run(prob, X, s, tau, bt, cuda_device, path)
    for b in range(B):
        for r in range(R):
            o = mo.log_posterior_and_grad_autograd(X[b, r], s[b, r], tau[b, r], bt[b, r], consts[b])
            assert abs(lp[b, r] - o[0]) <= TOL * abs(o[0])
            assert relerr(gX[b, r], o[1]) <= TOL
            assert relerr(gs[b, r], o[2]) <= TOL
            assert relerr(gt[b, r], o[3]) <= TOL


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("model", ["seir3", "seir4"])
def test_seir_vignette_shape(model, path, cuda_device):
    """n = 161, band 80: the reference's own configuration (vignette.ipynb:163)."""
    rng = np.random.default_rng(5)
    c, _, _ = seir_vignette_constants(0, model)
    prob = device_problem([c], model, cuda_device)
    X, s, tau = random_state(c, model, rng, 8, jitter=0.01)
    tau[:, :3] = np.array([6.0, 0.1, 1.6]) + 0.1 * rng.standard_normal((8, 3))
    bt = np.full((1, 8), 0.37)
    lp, gX, gs, gt = _run(prob, X[None], s[None], tau[None], bt, cuda_device, path)
    for r in (0, 3, 7):
        o = mo.log_posterior_and_grad_autograd(X[r], s[r], tau[r], 0.37, c)
        assert abs(lp[0, r] - o[0]) <= TOL * abs(o[0])
        assert relerr(gX[0, r], o[1]) <= TOL
        assert relerr(gs[0, r], o[2]) <= TOL
        assert relerr(gt[0, r], o[3]) <= TOL


@pytest.mark.parametrize("path", PATHS)
def test_committed_golden_vectors(path, cuda_device):
    """Inputs and expected outputs fixed in tests/golden/logpost_kat.npz (matrices from the genuine
    reference `_build_matrices`)."""
    import torch
    from magi_v2_b200 import ops
    g = load_golden("logpost_kat.npz")
    for name, model in mo.MODELS.items():
        ts, X_obs = g[f"{name}_ts"], g[f"{name}_X_obs"]
        mats = (g[f"{name}_Cinv"], g[f"{name}_m"], g[f"{name}_Kinv"])
        c = mo.make_constants(ts, X_obs, 1, g[f"{name}_phi1"], g[f"{name}_phi2"], int(g[f"{name}_band"]),
                              model.f_vec, matrices=mats)
        prob = device_problem([c], name, cuda_device)
        lp, gX, gs, gt = _run(prob, g[f"{name}_X"][None, None], g[f"{name}_s"][None, None],
                              g[f"{name}_tau"][None, None], np.full((1, 1), float(g[f"{name}_bt"])), cuda_device,
                              path)
        assert abs(lp[0, 0] - g[f"{name}_lp"]) <= TOL * abs(g[f"{name}_lp"])
        assert relerr(gX[0, 0], g[f"{name}_gX"]) <= TOL
        assert relerr(gs[0, 0], g[f"{name}_gs"]) <= TOL
        assert relerr(gt[0, 0], g[f"{name}_gt"]) <= TOL


@pytest.mark.parametrize("path", PATHS)
def test_linearity_in_temperature_and_chain_independence(path, cuda_device):
    """Size-independent properties: lp and gradient scale linearly with beta_temp (:348), and a chain's
    result does not depend on which other chains share its CTA."""
    rng = np.random.default_rng(7)
    c = synth_constants("seir4", seed=3, N=21, band=None)
    prob = device_problem([c], "seir4", cuda_device)
    X, s, tau = random_state(c, "seir4", rng, 8)
    a = _run(prob, X[None], s[None], tau[None], np.full((1, 8), 1.0), cuda_device, path)
    b = _run(prob, X[None], s[None], tau[None], np.full((1, 8), 0.25), cuda_device, path)
    for u, v in zip(a, b):
        assert relerr(0.25 * u, v) <= 1e-14
    perm = rng.permutation(8)
    p = _run(prob, X[None, perm], s[None, perm], tau[None, perm], np.full((1, 8), 1.0), cuda_device, path)
    for u, v in zip(a, p):
        assert np.array_equal(u[0][perm], v[0])


@pytest.mark.parametrize("model", ["seir4", "lorenz96"])
@pytest.mark.parametrize("path", PATHS)
def test_band_skipping_equals_dense(model, path, cuda_device):
    """Telling the kernels the bandsize (zero tiles are then not read) must not change the result:
    same banded matrices evaluated as banded and as dense (fast path and general path)."""
    rng = np.random.default_rng(21)
    c = synth_constants(model, seed=8, N=41, band=12)          # n = 81: 11 row blocks, band 12 -> kb = 2
    X, s, tau = random_state(c, model, rng, 8)
    bt = np.full((1, 8), 0.8)
    res = []
    for band in ("auto", None):
        prob = device_problem([c], model, cuda_device, band=band)
        assert prob.band == (12 if band == "auto" else -1)
        res.append(_run(prob, X[None], s[None], tau[None], bt, cuda_device, path))
    for u, v in zip(*res):
        assert relerr(u, v) <= 1e-13
    o = mo.log_posterior_and_grad_autograd(X[3], s[3], tau[3], 0.8, c)
    assert abs(res[0][0][0, 3] - o[0]) <= TOL * abs(o[0]) and relerr(res[0][1][0, 3], o[1]) <= TOL


@pytest.mark.parametrize("path", PATHS)
def test_sirw_n321_general_path(path, cuda_device):
    """Config 3 (SIRW, n = 321, D = 4, P = 5): np > 168, so the general path (vector arrays in the caller's
    workspace) runs.  Matrices from the oracle's reference route; parity to 1e-9."""
    rng = np.random.default_rng(33)
    c = synth_constants("sirw", seed=12, N=81, disc=2, band=None, T=4.0)      # 4*80 + 1 = 321 grid points
    assert c.n == 321
    prob = device_problem([c], "sirw", cuda_device)
    X, s, tau = random_state(c, "sirw", rng, 8, jitter=0.01)
    bt = np.full((1, 8), 0.5)
    lp, gX, gs, gt = _run(prob, X[None], s[None], tau[None], bt, cuda_device, path)
    for r in (0, 5):
        o = mo.log_posterior_and_grad_autograd(X[r], s[r], tau[r], 0.5, c)
        assert abs(lp[0, r] - o[0]) <= TOL * abs(o[0])
        assert relerr(gX[0, r], o[1]) <= TOL and relerr(gs[0, r], o[2]) <= TOL and relerr(gt[0, r], o[3]) <= TOL


@pytest.mark.parametrize("path", PATHS)
def test_lorenz96_large_grid_properties(path, cuda_device):
    """Config 5 shape (Lorenz-96, D = 10) at n = 641 with device-built matrices: too large for the oracle to
    finish in seconds, so the size-independent properties are checked -- linearity in beta_temp, independence
    of a chain from its CTA neighbours, and agreement of banded-as-banded with banded-as-dense."""
    import torch
    from magi_v2_b200 import ops
    rng = np.random.default_rng(9)
    n, D, R = 641, 10, 8
    I = np.linspace(0, 4, n)
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=cuda_device)
    phi1, phi2 = rng.uniform(0.5, 2.0, (1, D)), rng.uniform(0.15, 0.3, (1, D))
    C, Cp, Cpp = ops.cov_build(T(I), T(phi1), T(phi2), 2.01, True)
    Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, 160, 0.0)
    assert int(info.abs().max()) == 0
    packed = ops.pack_matrices(Cinv, m, Kinv)
    mask = np.zeros((1, n, D), dtype=np.uint8); mask[:, ::8] = 1
    y = rng.normal(2.0, 3.0, (1, n, D)) * mask
    consts = dict(mu=T(np.full((1, D), 2.0)), y=T(y), mask=T(mask, torch.uint8), N_ds=T(np.full((1, D), 81.0)),
                  beta=T(np.array([D * n / (81.0 * D)])), LB=T(np.full((1, D), 1e-4)), n=n)
    X = rng.normal(2.0, 3.0, (1, R, n, D)); s = rng.normal(-1, 0.5, (1, R, D)); tau = rng.normal(2.0, 0.2, (1, R, 1))
    res = {}
    for band in (160, None):
        prob = ops.PosteriorProblem("lorenz96", packed, band=band, **consts)
        res[band] = [a.cpu().numpy() for a in prob.logpost_grad(T(X), T(s), T(tau), T(np.full((1, R), 1.0)), path=path)]
        assert all(np.isfinite(a).all() for a in res[band])
    for u, v in zip(res[160], res[None]):
        assert relerr(u, v) <= 1e-12
    prob = ops.PosteriorProblem("lorenz96", packed, band=160, **consts)
    q = [a.cpu().numpy() for a in prob.logpost_grad(T(X), T(s), T(tau), T(np.full((1, R), 0.25)), path=path)]
    for u, v in zip(res[160], q):
        assert relerr(0.25 * u, v) <= 1e-14
    perm = rng.permutation(R)
    p = [a.cpu().numpy() for a in prob.logpost_grad(T(X[:, perm]), T(s[:, perm]), T(tau[:, perm]),
                                                    T(np.full((1, R), 1.0)), path=path)]
    for u, v in zip(res[160], p):   # (general path: chains 0-4 are summed by 3 warps, 5-7 by 2 -> not bit-identical)
        assert relerr(u[0][perm], v[0]) <= 1e-13
    # and against an independent float64 numpy evaluation of the same formula with the device-built matrices
    from oracle import magi_oracle as mo2
    idx = np.where(mask[0].reshape(-1) > 0)[0]
    oc = mo2.PosteriorConstants(I=I.reshape(-1, 1), mu_ds=np.full(D, 2.0), C_d_invs=Cinv[0].cpu().numpy(),
                                m_ds=m[0].cpu().numpy(), K_d_invs=Kinv[0].cpu().numpy(), N_ds=np.full(D, 81.0),
                                not_nan_idxs=idx, not_nan_cols=idx % D, y_tau_ds_observed=y[0].reshape(-1)[idx],
                                beta=float(D * n / (81.0 * D)), sigma_sqs_LB=np.full(D, 1e-4), f_vec=mo2.f_lorenz96)
    o = mo2.log_posterior_and_grad_autograd(X[0, 2], s[0, 2], tau[0, 2], 1.0, oc)
    assert abs(res[160][0][0, 2] - o[0]) <= TOL * abs(o[0]) and relerr(res[160][1][0, 2], o[1]) <= TOL


def test_wide_path_row_per_warp_split_matches_cta_path(cuda_device):
    """With enough block rows to fill the grid (here n = 641, D = 10, 48 chains: 81 * 10 * 6 block-row tasks) the
    wide path gives each warp a whole block row instead of splitting a row's tiles over the CTA's warps; the two
    grid shapes of the evaluation must still agree to rounding, and one chain is checked against the oracle."""
    import torch
    from magi_v2_b200 import ops
    rng = np.random.default_rng(19)
    n, D, R = 641, 10, 48
    I = np.linspace(0, 4, n)
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=cuda_device)
    phi1, phi2 = rng.uniform(0.5, 2.0, (1, D)), rng.uniform(0.15, 0.3, (1, D))
    C, Cp, Cpp = ops.cov_build(T(I), T(phi1), T(phi2), 2.01, True)
    Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, 160, 0.0)
    assert int(info.abs().max()) == 0
    mask = np.zeros((1, n, D), dtype=np.uint8); mask[:, ::8] = 1
    y = rng.normal(2.0, 3.0, (1, n, D)) * mask
    prob = ops.PosteriorProblem("lorenz96", ops.pack_matrices(Cinv, m, Kinv), mu=T(np.full((1, D), 2.0)), y=T(y),
                                mask=T(mask, torch.uint8), N_ds=T(np.full((1, D), 81.0)),
                                beta=T(np.array([D * n / (81.0 * D)])), LB=T(np.full((1, D), 1e-4)), n=n, band=160)
    X = rng.normal(2.0, 3.0, (1, R, n, D)); s = rng.normal(-1, 0.5, (1, R, D)); tau = rng.normal(2.0, 0.2, (1, R, 1))
    bt = rng.uniform(0.2, 1.2, (1, R))
    res = {p: [a.cpu().numpy() for a in prob.logpost_grad(T(X), T(s), T(tau), T(bt), path=p)] for p in PATHS}
    for u, v in zip(res["cta"], res["wide"]):
        assert np.isfinite(v).all() and relerr(v, u) <= 1e-12
    idx = np.where(mask[0].reshape(-1) > 0)[0]
    oc = mo.PosteriorConstants(I=I.reshape(-1, 1), mu_ds=np.full(D, 2.0), C_d_invs=Cinv[0].cpu().numpy(),
                               m_ds=m[0].cpu().numpy(), K_d_invs=Kinv[0].cpu().numpy(), N_ds=np.full(D, 81.0),
                               not_nan_idxs=idx, not_nan_cols=idx % D, y_tau_ds_observed=y[0].reshape(-1)[idx],
                               beta=float(D * n / (81.0 * D)), sigma_sqs_LB=np.full(D, 1e-4), f_vec=mo.f_lorenz96)
    r = 41
    o = mo.log_posterior_and_grad_autograd(X[0, r], s[0, r], tau[0, r], bt[0, r], oc)
    assert abs(res["wide"][0][0, r] - o[0]) <= TOL * abs(o[0]) and relerr(res["wide"][1][0, r], o[1]) <= TOL
    assert relerr(res["wide"][3][0, r], o[3]) <= TOL and relerr(res["wide"][2][0, r], o[2]) <= TOL


def test_wide_path_two_chain_groups_per_cta(cuda_device, monkeypatch):
    """The experimental MAGI_WIDE_NG=2 split of the wide path (two chain groups share every loaded matrix tile;
    slower at n = 1281, see posterior_wide.cu) returns what the default split returns, odd group counts included."""
    import torch
    from magi_v2_b200 import synth
    rng = np.random.default_rng(3)
    prob, info, state, _ = synth.sweep_problem(3, 20, cuda_device, seed0=5, model="seir4", bandsize=80)
    args = [torch.as_tensor(np.ascontiguousarray(state[k]), dtype=torch.float64, device=cuda_device)
            for k in ("X", "sig_pre", "th_pre")]
    bt = torch.as_tensor(rng.uniform(0.2, 1.0, (3, 20)), dtype=torch.float64, device=cuda_device)
    ref = [t.clone() for t in prob.logpost_grad(*args, bt, path="wide")]
    monkeypatch.setenv("MAGI_WIDE_NG", "2")
    got = prob.logpost_grad(*args, bt, path="wide")
    torch.cuda.synchronize()
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("N,disc,band", [(5, 1, 1), (5, 1, None), (9, 1, 2), (11, 2, 7), (12, 3, 16), (21, 3, 40),
                                         (21, 3, 9), (21, 3, None)])
def test_fast_path_edge_shapes(N, disc, band, cuda_device):
    """Grid sizes and bandwidths around the edges of the TMA-staged fast path (n = 9 ... 168 = its limit; one to 21
    block rows; bands of one tile, of every tile, not a multiple of 8; chain counts that leave a group partly empty):
    rings of one chunk per pass, rows shorter than a chunk, a last block row that is all padding but one point."""
    model, R, B = "seir4", 11, 2
    rng = np.random.default_rng(N * 100 + disc)
    consts = [synth_constants(model, seed=7 + b, N=N, disc=disc, band=band, T=2.0) for b in range(B)]
    n = consts[0].I.shape[0]
    assert n <= 168
    prob = device_problem(consts, model, cuda_device)
    st = [random_state(c, model, rng, R) for c in consts]
    X = np.stack([a[0] for a in st]); s = np.stack([a[1] for a in st]); tau = np.stack([a[2] for a in st])
    bt = rng.uniform(0.1, 1.5, (B, R))
    lp, gX, gs, gt = _run(prob, X, s, tau, bt, cuda_device, "cta")
    for b in range(B):
        for r in (0, 7, 8, 10):
            o = mo.log_posterior_and_grad_autograd(X[b, r], s[b, r], tau[b, r], bt[b, r], consts[b])
            assert abs(lp[b, r] - o[0]) <= TOL * abs(o[0])
            assert relerr(gX[b, r], o[1]) <= TOL
            assert relerr(gs[b, r], o[2]) <= TOL
            assert relerr(gt[b, r], o[3]) <= TOL


def test_fast_path_sirw_at_the_reference_grid(cuda_device):
    """SIRW (D = 4, P = 5) at n = 161, band 80: the larger per-chain scalar area leaves the tile rings less room than
    SEIR4 has, so the ring plan caps the chunk size and the long rows take four chunks per pass instead of three."""
    model, R = "sirw", 8
    rng = np.random.default_rng(42)
    c = synth_constants(model, seed=3, N=21, disc=3, band=80, T=4.0, nan_frac=0.3)
    assert c.I.shape[0] == 161
    prob = device_problem([c], model, cuda_device)
    X, s, tau = random_state(c, model, rng, R)
    bt = rng.uniform(0.2, 1.2, (1, R))
    lp, gX, gs, gt = _run(prob, X[None], s[None], tau[None], bt, cuda_device, "cta")
    for r in (0, 5):
        o = mo.log_posterior_and_grad_autograd(X[r], s[r], tau[r], bt[0, r], c)
        assert abs(lp[0, r] - o[0]) <= TOL * abs(o[0])
        assert relerr(gX[0, r], o[1]) <= TOL
        assert relerr(gs[0, r], o[2]) <= TOL
        assert relerr(gt[0, r], o[3]) <= TOL
