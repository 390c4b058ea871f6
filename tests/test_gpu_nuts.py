"""NUTS on the CUDA log-posterior + gradient operator.

(1) draw-for-draw: the batched tree builder of `magi_v2_b200/nuts.py` driving `magi_b200_logpost_grad` builds the
same trees, takes the same number of leapfrogs and lands on the same states as the recursive single-chain oracle
(`oracle.nuts_chain`) driving the oracle's log-posterior, for several transitions with dual averaging and tempering.
(2) the API: `MAGI_v2.predict(sampler="nuts")` and the batched front end run end to end."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests.helpers import device_problem, load_golden, random_state, relerr, synth_constants

pytestmark = pytest.mark.gpu


def _T(a, device):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)


@pytest.mark.parametrize("model,fused", [("seir3", True), ("sirw", True), ("seir3", False), ("lorenz96", True)])
def test_nuts_chain_matches_oracle_draw_for_draw(model, fused, cuda_device):
    import torch
    from magi_v2_b200 import nuts
    rng = np.random.default_rng(7)
    B, R, n_iter, max_depth = 2, 3, 5, 5
    consts = [synth_constants(model, seed=80 + b, N=9) for b in range(B)]
    prob = device_problem(consts, model, cuda_device)
    n, D, P = consts[0].n, prob.D, prob.P
    st = [random_state(c, model, rng, R, jitter=0.005) for c in consts]
    X = np.stack([a[0] for a in st]); s = np.stack([a[1] for a in st]); tau = np.stack([a[2] for a in st])
    eps0, seed, num_adapt = 4e-4, 99, 4
    z = nuts.pack_state(_T(X, cuda_device), _T(s, cuda_device), _T(tau, cuda_device))
    eps = torch.full((B * R,), eps0, dtype=torch.float64, device=cuda_device)
    da = torch.zeros((B * R, 4), dtype=torch.float64, device=cuda_device)
    da[:, 2] = float(np.log(10.0 * eps0))
    keep = []
    # fused: the product path (CUDA bookkeeping kernels); not fused: the tensor-op form of the same loop body
    out = nuts.nuts_run_(z, eps, da, nuts.problem_value_and_grad(prob, R), n_iter=n_iter, num_adapt=num_adapt,
                         seed=seed, max_tree_depth=max_depth, on_sample=lambda it, zz, info: keep.append(zz.cpu().numpy()),
                         leaf_engine=nuts.FusedLeafEngine(prob, R) if fused else None)
    torch.cuda.synchronize()
    nl = out["n_leapfrog"].cpu().numpy()
    acc = out["accept_prob"].cpu().numpy()
    assert nl.max() > 1
    for b in range(B):
        for r in range(R):
            c = b * R + r
            zs, accs, epss, nls = mo.nuts_chain(consts[b], model, mo.pack_state(X[b, r], s[b, r], tau[b, r]), n_iter,
                                                eps0, seed, c, num_adaptation_steps=num_adapt, max_tree_depth=max_depth)
            assert np.array_equal(nl[:, c], nls), (nl[:, c], nls)
            assert np.allclose(acc[:, c], accs, rtol=0, atol=1e-7)
            for it in range(n_iter):
                assert relerr(keep[it][c], zs[it]) <= 1e-8


def test_predict_with_nuts_runs_the_reference_sampler_stack(cuda_device):
    """vignette.ipynb configuration (SEIR3, n = 161, band 80) with the reference's sampler: short run, finite samples,
    positive parameters, trees deeper than one leapfrog, acceptance statistic driven towards the 0.75 target."""
    from magi_v2_b200 import MAGI_v2
    g = load_golden("seir_datasets.npz")
    X = g["X_obs"][0][:, 1:].copy()
    X[X < 0.0] = 0.0
    model = MAGI_v2(D_thetas=3, ts_obs=g["ts_obs"], X_obs=X, bandsize=80, f_vec="seir3")
    model.initial_fit(discretization=1, verbose=False)
    res = model.predict(num_results=40, num_burnin_steps=60, sampler="nuts", n_chains=4, seed=1, max_tree_depth=6)
    th = res["thetas_samps"]
    assert th.shape == (4, 40, 3) and res["X_samps"].shape == (4, 40, 161, 3)
    assert np.isfinite(th).all() and np.all(th > 0) and np.all(res["sigma_sqs_samps"] > 0)
    kr = res["kernel_results"]
    assert kr["sampler"] == "nuts" and kr["leapfrogs_taken"].max() > 1
    assert 0.3 < kr["accept_prob"].mean() <= 1.0
    assert np.all(kr["step_size"] > 0) and not np.allclose(kr["step_size"], 0.1)


def test_magi_batch_front_end(cuda_device):
    """BASELINE config 2 through `MagiBatch` (the batched `initial_fit` + `predict`): 20 datasets, one launch."""
    from magi_v2_b200.batch import MagiBatch
    g = load_golden("seir_datasets.npz")
    X = g["X_obs"][1:].copy()
    X[X < 0.0] = 0.0
    mb = MagiBatch(g["ts_obs"], X, 80, "seir4", device=cuda_device).initial_fit(1, hparam_iters=200)
    assert mb.thetas_init.shape == (20, 3) and np.all(np.isfinite(mb.thetas_init))
    res = mb.predict(num_results=200, num_burnin_steps=300, n_chains=4, n_leapfrog=16, seed=2, sampler="hmc")
    assert res["thetas_samps"].shape == (20, 4, 200, 3) and res["sigma_sqs_samps"].shape == (20, 4, 200, 4)
    assert np.isfinite(res["thetas_samps"]).all() and np.all(res["thetas_samps"] > 0)
    assert res["X_mean"].shape == (20, 161, 4) and np.isfinite(res["X_sd"]).all()
    assert res["accept_prob"].mean() > 0.3
    rn = mb.predict(num_results=10, num_burnin_steps=20, n_chains=2, seed=2, sampler="nuts", max_tree_depth=5)
    assert rn["thetas_samps"].shape == (20, 2, 10, 3) and np.isfinite(rn["thetas_samps"]).all()
    assert rn["leapfrogs_taken"].max() > 1 and rn["X_mean"].shape == (20, 161, 4)


def test_nuts_posterior_means_within_monte_carlo_error(cuda_device):
    """north_star: 'posterior means of theta within Monte Carlo standard error' -- for the reference's sampler.  64 NUTS
    chains on the CUDA operator and 8 independently seeded fixed-length HMC chains of the oracle (CPU) sample the same
    untempered posterior of a small SEIR3 problem; theta and log sigma^2 means agree within 5 combined standard errors."""
    import torch
    from magi_v2_b200 import nuts
    model = "seir3"
    c = synth_constants(model, seed=77, N=9, nan_frac=0.0)
    prob = device_problem([c], model, cuda_device)
    n, D, P = c.n, prob.D, prob.P
    rng = np.random.default_rng(4)
    R, burn, keep, eps0 = 64, 200, 400, 2e-3
    X, s, tau = random_state(c, model, rng, R, jitter=0.005)
    s[:] = -5.0 + 0.1 * rng.standard_normal(s.shape)
    tau[:] = 0.5 + 0.1 * rng.standard_normal(tau.shape)
    z = nuts.pack_state(_T(X[None], cuda_device), _T(s[None], cuda_device), _T(tau[None], cuda_device))
    eps = torch.full((R,), eps0, dtype=torch.float64, device=cuda_device)
    da = torch.zeros((R, 4), dtype=torch.float64, device=cuda_device)
    da[:, 2] = float(np.log(10.0 * eps0))
    eng = nuts.FusedLeafEngine(prob, R)
    nuts.nuts_run_(z, eps, da, None, n_iter=burn, num_adapt=160, seed=21, fixed_beta_temp=1.0, max_tree_depth=8,
                   leaf_engine=eng)
    kept = []
    out = nuts.nuts_run_(z, eps, da, None, n_iter=keep, iter0=burn, num_adapt=160, seed=21, fixed_beta_temp=1.0,
                         max_tree_depth=8, leaf_engine=eng, on_sample=lambda it, zz, info: kept.append(zz[:, n * D:].clone()))
    torch.cuda.synchronize()
    tail = torch.stack(kept).cpu().numpy()                       # [keep, R, D + P]
    assert 0.55 < float(out["accept_prob"].mean()) < 0.95 and float(out["n_leapfrog"].double().mean()) > 3
    th_g = np.logaddexp(0, tail[:, :, D:])
    ls_g = np.log(np.logaddexp(0, tail[:, :, :D]) + c.sigma_sqs_LB)
    Rc, burn_c, keep_c, L = 8, 200, 400, 16
    th_c, ls_c = [], []
    for r in range(Rc):
        zs, _, _, _ = mo.hmc_chain(c, model, mo.pack_state(X[r], s[r], tau[r]), burn_c + keep_c, L, eps0, seed=999,
                                   chain_id=r, num_adaptation_steps=160, fixed_beta_temp=1.0)
        zs = np.array(zs[burn_c:])
        th_c.append(np.logaddexp(0, zs[:, n * D + D:]))
        ls_c.append(np.log(np.logaddexp(0, zs[:, n * D:n * D + D]) + c.sigma_sqs_LB))
    for g, cc, nm in ((th_g, np.array(th_c), "theta"), (ls_g, np.array(ls_c), "log sigma^2")):
        mg, mc = g.mean(axis=0), cc.mean(axis=1)                  # per-chain means [R, k], [Rc, k]
        se = np.sqrt(mg.var(axis=0, ddof=1) / R + mc.var(axis=0, ddof=1) / Rc)
        zsc = (mg.mean(axis=0) - mc.mean(axis=0)) / se
        print(nm, "nuts (cuda)", mg.mean(axis=0), "hmc (oracle)", mc.mean(axis=0), "z", zsc)
        assert np.all(np.abs(zsc) < 5.0), (nm, zsc)
