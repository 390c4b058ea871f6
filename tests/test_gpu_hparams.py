"""GP hyper-parameter fit (`magi_v2.py:538-691`): the closed-form gradient used
instead of TF autodiff is checked against central differences of the objective, and the fit is checked to
increase the objective and to land on sensible values for GP-generated data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(cuda_device, n=41, B=2, D=3, seed=0):
    import torch
    from magi_v2_b200 import hparams
    rng = np.random.default_rng(seed)
    I = np.linspace(0, 4, n)
    X = np.stack([np.stack([np.sin((1 + d) * I + b) * (0.2 + 0.1 * d) + 0.02 * rng.standard_normal(n)
                            for d in range(D)], axis=1) for b in range(B)])           # [B,n,D]
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda_device)
    mu_phi2, sd_phi2 = hparams.fourier_prior(X)
    sd = X.std(axis=1)
    loc = T(np.stack([np.full((B, D), 1e-4), mu_phi2, (0.1 * sd) ** 2]))
    scale = T(np.stack([np.full((B, D), 1000.0 * np.sqrt(D)), sd_phi2 * np.sqrt(D), np.full((B, D), 1000.0 * np.sqrt(D))]))
    grid = T(I)
    dt = grid[:, None] - grid[None, :]
    x = T(np.transpose(X, (0, 2, 1)))
    xc = x - x.mean(dim=-1, keepdim=True)
    v = T(np.log(np.expm1(np.stack([sd ** 2, mu_phi2, (0.1 * sd) ** 2]))))
    return hparams, I, X, v, grid, dt, xc, loc, scale


def test_closed_form_gradient_matches_central_differences(cuda_device):
    hparams, I, X, v, grid, dt, xc, loc, scale = _setup(cuda_device)
    obj, g = hparams.objective_and_grad(v, grid, dt, xc, loc, scale)
    g = g.cpu().numpy()
    h = 1e-5
    for k in range(3):
        for b in range(v.shape[1]):
            for d in range(v.shape[2]):
                vp, vm = v.clone(), v.clone()
                vp[k, b, d] += h
                vm[k, b, d] -= h
                op, _ = hparams.objective_and_grad(vp, grid, dt, xc, loc, scale)
                om, _ = hparams.objective_and_grad(vm, grid, dt, xc, loc, scale)
                fd = -float(op[b, d] - om[b, d]) / (2 * h)          # g is the gradient of the LOSS = -objective
                assert abs(fd - g[k, b, d]) <= 1e-5 * max(1.0, abs(fd)), (k, b, d, fd, g[k, b, d])


def test_fit_improves_the_objective_and_recovers_the_noise_level(cuda_device):
    hparams, I, X, v, grid, dt, xc, loc, scale = _setup(cuda_device, n=81, B=1, D=2, seed=3)
    obj0, _ = hparams.objective_and_grad(v, grid, dt, xc, loc, scale)
    hp = hparams.fit_kernel_hparams(I, X, device=cuda_device, num_iters=400)
    import torch
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda_device)
    v1 = T(np.log(np.expm1(np.stack([hp["phi1s"], hp["phi2s"], hp["sigma_sqs"]]))))
    obj1, _ = hparams.objective_and_grad(v1, grid, dt, xc, loc, scale)
    assert np.all(obj1.cpu().numpy() > obj0.cpu().numpy())
    assert np.all(hp["phi1s"] > 0) and np.all(hp["phi2s"] > 0.05) and np.all(hp["phi2s"] < 3.0)
    assert np.all(hp["sigma_sqs"] < 10 * 0.02 ** 2)               # generating noise sd was 0.02


def test_objective_gradient_and_adam_steps_match_the_oracle_restatement(cuda_device):
    """Row f1 against the oracle (oracle/init_oracle.py, an autograd restatement of magi_v2.py:574-653 with the (D, D)
    broadcast, truncated-normal priors, softplus variables, jitter 1e-6) on the vignette data: the objective and its
    gradient at the reference's start values, and 25 Adam steps of the fit (tests/golden/init_kat.npz)."""
    import torch
    from magi_v2_b200 import hparams
    from tests.helpers import load_golden
    g = load_golden("init_kat.npz")
    I, Xi = g["I"].ravel(), g["X_interp"]
    D = Xi.shape[1]
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda_device)
    mu_phi2, sd_phi2 = hparams.fourier_prior(Xi[None])
    sd = Xi.std(axis=0)[None]
    loc = T(np.stack([np.full((1, D), 1e-4), mu_phi2, (0.1 * sd) ** 2]))
    scale = T(np.stack([np.full((1, D), 1000.0 * np.sqrt(D)), sd_phi2 * np.sqrt(D), np.full((1, D), 1000.0 * np.sqrt(D))]))
    grid = T(I)
    dt = grid[:, None] - grid[None, :]
    x = T(Xi.T[None])
    xc = x - x.mean(dim=-1, keepdim=True)
    v = T(g["hp_v0"][:, None, :])                                            # (phi1, phi2, sigma^2) pre-activations
    obj, grad = hparams.objective_and_grad(v, grid, dt, xc, loc, scale, True)
    # oracle loss [D, D]: element (i, j) = -(priors_j + gp_i); its diagonal is -(objective of component i) up to the
    # constant normalisers of the truncated normals, so compare differences between components of the GP part via the
    # gradient, which has no constants: d(sum of the [D, D] loss) = D x the per-component gradient
    assert np.allclose(D * grad[:, 0].cpu().numpy(), g["hp_grad0"], rtol=1e-7, atol=1e-9)
    hp = hparams.fit_kernel_hparams(I, Xi[None], device=cuda_device, num_iters=25)
    sp = lambda a: np.logaddexp(0.0, a)
    last = g["hp_trace25"][-1]                                               # [3, D] pre-activations after 25 steps
    for k, name in enumerate(("phi1s", "phi2s", "sigma_sqs")):
        assert np.allclose(hp[name][0], sp(last[k]), rtol=1e-7), name
