"""Parity of the CUDA log-posterior + gradient (C-ABI magi_b200_logpost_grad) with the oracle's
op-for-op restatement of magi_v2.py:308-348 + autograd.  Tolerance: 1e-9 relative (BASELINE.json
north_star), FP64, given identical C^-1, m, K^-1."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests.helpers import device_problem, load_golden, random_state, relerr, seir_vignette_constants, synth_constants

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _run(prob, X, s, tau, bt, device):
    import torch
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
    lp, gX, gs, gt = prob.logpost_grad(T(X), T(s), T(tau), T(bt))
    torch.cuda.synchronize()
    return lp.cpu().numpy(), gX.cpu().numpy(), gs.cpu().numpy(), gt.cpu().numpy()


@pytest.mark.parametrize("model", ["seir3", "seir4", "sirw", "lorenz96"])
@pytest.mark.parametrize("R", [1, 8, 11])
def test_small_batches_match_oracle(model, R, cuda_device):
    B = 3
    rng = np.random.default_rng(10 + R)
    consts = [synth_constants(model, seed=100 + b) for b in range(B)]
    prob = device_problem(consts, model, cuda_device)
    st = [random_state(c, model, rng, R) for c in consts]
    X = np.stack([a[0] for a in st]); s = np.stack([a[1] for a in st]); tau = np.stack([a[2] for a in st])
    bt = rng.uniform(0.1, 1.5, (B, R))
    lp, gX, gs, gt = _run(prob, X, s, tau, bt, cuda_device)
    for b in range(B):
        for r in range(R):
            o = mo.log_posterior_and_grad_autograd(X[b, r], s[b, r], tau[b, r], bt[b, r], consts[b])
            assert abs(lp[b, r] - o[0]) <= TOL * abs(o[0])
            assert relerr(gX[b, r], o[1]) <= TOL
            assert relerr(gs[b, r], o[2]) <= TOL
            assert relerr(gt[b, r], o[3]) <= TOL


@pytest.mark.parametrize("model", ["seir3", "seir4"])
def test_seir_vignette_shape(model, cuda_device):
    """n = 161, band 80: the reference's own configuration (vignette.ipynb:163)."""
    rng = np.random.default_rng(5)
    c, _, _ = seir_vignette_constants(0, model)
    prob = device_problem([c], model, cuda_device)
    X, s, tau = random_state(c, model, rng, 8, jitter=0.01)
    tau[:, :3] = np.array([6.0, 0.1, 1.6]) + 0.1 * rng.standard_normal((8, 3))
    bt = np.full((1, 8), 0.37)
    lp, gX, gs, gt = _run(prob, X[None], s[None], tau[None], bt, cuda_device)
    for r in (0, 3, 7):
        o = mo.log_posterior_and_grad_autograd(X[r], s[r], tau[r], 0.37, c)
        assert abs(lp[0, r] - o[0]) <= TOL * abs(o[0])
        assert relerr(gX[0, r], o[1]) <= TOL
        assert relerr(gs[0, r], o[2]) <= TOL
        assert relerr(gt[0, r], o[3]) <= TOL


def test_committed_golden_vectors(cuda_device):
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
                              g[f"{name}_tau"][None, None], np.full((1, 1), float(g[f"{name}_bt"])), cuda_device)
        assert abs(lp[0, 0] - g[f"{name}_lp"]) <= TOL * abs(g[f"{name}_lp"])
        assert relerr(gX[0, 0], g[f"{name}_gX"]) <= TOL
        assert relerr(gs[0, 0], g[f"{name}_gs"]) <= TOL
        assert relerr(gt[0, 0], g[f"{name}_gt"]) <= TOL


def test_linearity_in_temperature_and_chain_independence(cuda_device):
    """Size-independent properties: lp and gradient scale linearly with beta_temp (:348), and a chain's
    result does not depend on which other chains share its CTA."""
    rng = np.random.default_rng(7)
    c = synth_constants("seir4", seed=3, N=21, band=None)
    prob = device_problem([c], "seir4", cuda_device)
    X, s, tau = random_state(c, "seir4", rng, 8)
    a = _run(prob, X[None], s[None], tau[None], np.full((1, 8), 1.0), cuda_device)
    b = _run(prob, X[None], s[None], tau[None], np.full((1, 8), 0.25), cuda_device)
    for u, v in zip(a, b):
        assert relerr(0.25 * u, v) <= 1e-14
    perm = rng.permutation(8)
    p = _run(prob, X[None, perm], s[None, perm], tau[None, perm], np.full((1, 8), 1.0), cuda_device)
    for u, v in zip(a, p):
        assert np.array_equal(u[0][perm], v[0])


@pytest.mark.parametrize("model", ["seir4", "lorenz96"])
def test_band_skipping_equals_dense(model, cuda_device):
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
        res.append(_run(prob, X[None], s[None], tau[None], bt, cuda_device))
    for u, v in zip(*res):
        assert relerr(u, v) <= 1e-13
    o = mo.log_posterior_and_grad_autograd(X[3], s[3], tau[3], 0.8, c)
    assert abs(res[0][0][0, 3] - o[0]) <= TOL * abs(o[0]) and relerr(res[0][1][0, 3], o[1]) <= TOL
