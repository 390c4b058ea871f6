"""A user-supplied f_vec end to end: traced (tests/test_tracing_cpu.py), compiled at run
time into its own library, evaluated by `magi_b200_logpost_grad_wide` for model id MAGI_MODEL_USER, sampled by the
reference's sampler stack."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests.helpers import device_problem, load_golden, random_state, relerr, synth_constants
from tests.test_tracing_cpu import saturating_sir, seir_tf_style

pytestmark = pytest.mark.gpu


def _T(a, device):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)


def test_compiled_user_system_matches_the_compiled_in_one(cuda_device):
    """The TF-style SEIR callable, traced and compiled, against the registry's seir3 on the same inputs: <= 1e-12."""
    import torch
    from magi_v2_b200 import models, ops
    um = models.resolve(seir_tf_style, 3, 3, jit=True)
    assert um.lib_path and um.model_id == 100
    rng = np.random.default_rng(4)
    consts = [synth_constants("seir3", seed=90 + b, N=21) for b in range(2)]
    ref = device_problem(consts, "seir3", cuda_device)
    usr = ops.PosteriorProblem(um, ref.packed, ref.mu, ref.y, ref.mask, ref.N_ds, ref.beta, ref.LB, n=ref.n, band=ref.band)
    st = [random_state(c, "seir3", rng, 5) for c in consts]
    X, s, tau = (_T(np.stack([a[k] for a in st]), cuda_device) for k in range(3))
    bt = _T(rng.uniform(0.2, 1.4, (2, 5)), cuda_device)
    a = ref.logpost_grad(X, s, tau, bt, path="wide")
    b = usr.logpost_grad(X, s, tau, bt)
    torch.cuda.synchronize()
    assert usr.eval_path(5) == "wide"
    for u, v in zip(a, b):
        assert relerr(v.cpu().numpy(), u.cpu().numpy()) <= 1e-12
    with pytest.raises(ValueError):
        usr.eval_path(5, "cta")


def test_nonlinear_user_system_matches_oracle_autograd(cuda_device):
    """A system that is not affine in theta and uses exp(): the compiled operator against the oracle's autograd
    restatement of magi_v2.py:308-348 with the same callable (torch ops), 1e-9."""
    import torch
    from magi_v2_b200 import models, ops

    def f_torch(t, X, th):
        S, I = X[:, 0:1], X[:, 1:2]
        inc = th[0] * S * I / (1.0 + th[2] * I)
        return torch.cat([-inc, inc - th[1] * torch.exp(-0.1 * I) * I], dim=1)

    um = models.resolve(saturating_sir, 2, 3, jit=True)
    assert not um.affine_in_theta
    rng = np.random.default_rng(6)
    N, band = 17, 8
    ts = np.linspace(0, 2, N)
    X_obs = np.abs(rng.normal(0.3, 0.1, (N, 2)))
    X_obs[rng.uniform(size=X_obs.shape) < 0.2] = np.nan
    X_obs[0] = 0.3; X_obs[-1] = 0.2
    c = mo.make_constants(ts, X_obs, 1, rng.uniform(0.01, 0.05, 2), rng.uniform(0.2, 0.5, 2), band, f_torch)
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=cuda_device)
    packed = ops.pack_matrices(T(c.C_d_invs[None]), T(c.m_ds[None]), T(c.K_d_invs[None]))
    y, mask = c.dense_y_mask()
    prob = ops.PosteriorProblem(um, packed, mu=T(c.mu_ds[None]), y=T(y[None]), mask=T(mask[None], torch.uint8),
                                N_ds=T(c.N_ds[None].astype(np.float64)), beta=T(np.array([c.beta])),
                                LB=T(c.sigma_sqs_LB[None]), n=c.n, band=band)
    R = 3
    Xs = np.where(mask > 0, y, 0.3)[None] + 0.02 * rng.standard_normal((R, c.n, 2))
    s, tau, bt = rng.normal(-4, 1, (R, 2)), rng.normal(0.3, 0.5, (R, 3)), rng.uniform(0.3, 1.2, R)
    lp, gX, gs, gt = prob.logpost_grad(T(Xs[None]), T(s[None]), T(tau[None]), T(bt[None]))
    torch.cuda.synchronize()
    for r in range(R):
        o = mo.log_posterior_and_grad_autograd(Xs[r], s[r], tau[r], bt[r], c)
        assert abs(float(lp[0, r]) - o[0]) <= 1e-9 * abs(o[0])
        assert relerr(gX[0, r].cpu().numpy(), o[1]) <= 1e-9
        assert relerr(gs[0, r].cpu().numpy(), o[2]) <= 1e-9 and relerr(gt[0, r].cpu().numpy(), o[3]) <= 1e-9


def test_magi_v2_accepts_a_tf_style_callable(cuda_device):
    """The drop-in class with a user callable (jit=True forces the traced path although the callable happens to
    reproduce a compiled-in system): initial fit and a short run of the reference's sampler."""
    from magi_v2_b200 import MAGI_v2
    g, f = load_golden("seir_datasets.npz"), load_golden("vignette_fit.npz")
    X = g["X_obs"][0][:, 1:].copy()
    X[X < 0.0] = 0.0
    hp = {"phi1s": f["phi1s"], "phi2s": f["phi2s"], "sigma_sqs": f["sigma_sqs"]}
    model = MAGI_v2(D_thetas=3, ts_obs=g["ts_obs"], X_obs=X, bandsize=80, f_vec=seir_tf_style, jit=True)
    assert model.model.lib_path is not None
    model.initial_fit(discretization=1, hparams=hp)
    assert np.allclose(model.thetas_init, f["thetas_init_reference"], rtol=2e-2, atol=2e-3)
    res = model.predict(num_results=5, num_burnin_steps=5, n_chains=2, seed=3, max_tree_depth=5)
    ref = MAGI_v2(D_thetas=3, ts_obs=g["ts_obs"], X_obs=X, bandsize=80, f_vec="seir3")
    ref.initial_fit(discretization=1, hparams=hp)
    rr = ref.predict(num_results=5, num_burnin_steps=5, n_chains=2, seed=3, max_tree_depth=5)
    assert res["thetas_samps"].shape == (2, 5, 3) and np.isfinite(res["X_samps"]).all()
    # same seed, same trees: the traced system and the compiled-in one give the same chains
    assert np.array_equal(res["kernel_results"]["leapfrogs_taken"], rr["kernel_results"]["leapfrogs_taken"])
    assert relerr(res["thetas_samps"], rr["thetas_samps"]) <= 1e-8
