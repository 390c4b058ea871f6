"""Gradient-matching initialisation of a completely unobserved component (magi_v2.py:182-250) -- host logic."""
import numpy as np

from magi_v2_b200 import init_fit, models
from tests.helpers import load_golden


def test_registry_rhs_agree_between_numpy_and_torch():
    import torch
    rng = np.random.default_rng(0)
    for m in models.REGISTRY.values():
        X, th = rng.uniform(0.05, 0.6, (9, m.D)), rng.uniform(0.1, 2.0, m.P)
        a = m.f_vec(None, X, th)
        b = m.f_vec(None, torch.as_tensor(X), torch.as_tensor(th)).numpy()
        np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-15)


def test_unobserved_S_of_seir4_is_recovered_by_gradient_matching():
    g = load_golden("seir_datasets.npz")
    ts, Xt = g["ts_obs"], g["X_true"][0]                   # noise-free truth on the observation grid [N, 4]
    n = 2 * (len(ts) - 1) + 1
    I = np.interp(np.arange(n), np.arange(n)[::2], ts)
    Xd = np.stack([np.interp(I, ts, Xt[:, d]) for d in range(4)], axis=1)
    m = models.REGISTRY["seir4"]
    Xu, th, l0, l1 = init_fit.fit_unobserved(m, I, Xd[:, 1:], [1, 2, 3], [0], Xd[:, 1:], num_iters=4000, seed=0)
    assert Xu.shape == (n, 1) and th.shape == (3,) and np.isfinite(Xu).all() and np.isfinite(th).all()
    assert l1 < 1e-2 * l0
    # gamma and sigma are identified by the observed E, I, R equations alone: (0.6, 1.8)
    assert abs(th[1] - 0.6) < 0.1 and abs(th[2] - 1.8) < 0.3, th
