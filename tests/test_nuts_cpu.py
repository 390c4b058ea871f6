"""Host logic of the batched NUTS driver (magi_v2_b200/nuts.py) against the recursive single-chain oracle
(oracle/magi_oracle.py::nuts_transition) on a toy target, draw for draw: same Philox counters, same tree, same
proposal.  The product wires `value_and_grad` to the CUDA operator (tests/test_gpu_nuts.py); here it is a closed-form
target evaluated with torch on the CPU so that the tree logic itself is covered without a GPU."""
import math

import numpy as np
import pytest
import torch

from magi_v2_b200 import nuts
from oracle import magi_oracle as orc


def test_philox_matches_oracle_stream():
    ids = torch.tensor([0, 5, 77], dtype=torch.int64)
    z = nuts.rng_normals(1234567890123, ids, 9, 11).numpy()
    for r, c in enumerate([0, 5, 77]):
        np.testing.assert_allclose(z[r], orc.rng_normals(1234567890123, c, 9, 11), rtol=0, atol=1e-14)
    a, b = nuts.rng_uniform_pairs(42, ids, 3, nuts.RNG_NUTS_LEAF, 7, 4)
    for r, c in enumerate([0, 5, 77]):
        for k in range(4):
            ua, ub = orc.rng_uniform_pair(42, c, 3, orc.RNG_PURPOSE_NUTS_LEAF, 7 + k)
            assert a[r, k].item() == ua and b[r, k].item() == ub


def _target(S, seed):
    """Correlated Gaussian + a quartic term so that trees differ between chains."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((S, S))
    Q = A @ A.T / S + np.diag(np.linspace(0.5, 30.0, S))
    Qt = torch.as_tensor(Q)

    def vg_np(z):
        return float(-0.5 * z @ Q @ z - 0.05 * np.sum(z ** 4)), -(Q @ z) - 0.2 * z ** 3

    def vg_t(z):
        return -0.5 * torch.einsum("cs,st,ct->c", z, Qt, z) - 0.05 * (z ** 4).sum(1), -(z @ Qt) - 0.2 * z ** 3

    return vg_np, vg_t


@pytest.mark.parametrize("eps0,max_depth", [(0.02, 6), (0.15, 10), (0.9, 10)])
def test_transition_matches_recursive_oracle(eps0, max_depth):
    S, C, seed = 7, 12, 2024
    vg_np, vg_t = _target(S, 1)
    rng = np.random.default_rng(3)
    z0 = rng.standard_normal((C, S)) * 0.4
    eps = eps0 * (1.0 + 0.5 * rng.random(C))
    ids = torch.arange(100, 100 + C, dtype=torch.int64)
    depths = []
    for it in range(3):
        z = torch.as_tensor(z0.copy())
        info = nuts.nuts_transition(z, torch.as_tensor(eps), vg_t, seed, ids, it, max_tree_depth=max_depth, sync_every=2)
        for c in range(C):
            zn, lpn, acc, nl, depth, div = orc.nuts_transition(z0[c], eps[c], vg_np, seed, 100 + c, it, max_depth)
            assert int(info["n_leapfrog"][c]) == nl and int(info["depth"][c]) == depth
            assert bool(info["diverged"][c]) == div
            np.testing.assert_allclose(z[c].numpy(), zn, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(info["lp"][c].item(), lpn, rtol=1e-10)
            np.testing.assert_allclose(info["accept_stat"][c].item(), acc, rtol=1e-10)
            depths.append(depth)
        z0 = z.numpy().copy()
    assert len(set(depths)) > 1 or max_depth == 6          # the chains really build different trees


def test_dual_averaging_matches_oracle():
    C, num_adapt = 4, 7
    rng = np.random.default_rng(0)
    eps = torch.full((C,), 0.1, dtype=torch.float64)
    da = torch.zeros((C, 4), dtype=torch.float64)
    da[:, 2] = math.log(10 * 0.1)
    sts = [orc.DualAveragingState.create(0.1) for _ in range(C)]
    for it in range(10):
        a = rng.random(C)
        nuts.dual_averaging_update_(eps, da, torch.as_tensor(a), num_adapt)
        for c in range(C):
            sts[c] = orc.dual_averaging_update(sts[c], a[c], num_adapt)
            np.testing.assert_allclose(eps[c].item(), sts[c].step_size, rtol=1e-13)


def test_nuts_samples_a_gaussian():
    """Stationarity: mean and variance of N(mu, diag(s^2)) recovered within Monte-Carlo error."""
    S, C = 5, 64
    mu = torch.linspace(-1, 1, S, dtype=torch.float64)
    sd = torch.linspace(0.5, 2.0, S, dtype=torch.float64)

    def vg_at(z, bt):
        d = (z - mu) / sd
        return -0.5 * (d * d).sum(1), -d / sd

    z = torch.zeros((C, S), dtype=torch.float64)
    eps = torch.full((C,), 0.1, dtype=torch.float64)
    da = torch.zeros((C, 4), dtype=torch.float64); da[:, 2] = math.log(1.0)
    nuts.nuts_run_(z, eps, da, vg_at, n_iter=150, num_adapt=120, fixed_beta_temp=1.0, seed=5, max_tree_depth=6)
    keep = []
    out = nuts.nuts_run_(z, eps, da, vg_at, n_iter=300, iter0=150, num_adapt=120, fixed_beta_temp=1.0, seed=5,
                         max_tree_depth=6, on_sample=lambda it, zz, info: keep.append(zz.clone()))
    s = torch.stack(keep)                                   # [300, C, S]
    assert 0.6 < out["accept_prob"].mean().item() < 0.95
    m, v = s.mean((0, 1)), s.var((0, 1))
    assert torch.all((m - mu).abs() < 0.08 * sd * 2), (m - mu)
    assert torch.all((v / sd ** 2 - 1).abs() < 0.08), v / sd ** 2


def test_results_do_not_depend_on_how_chains_are_sharded():
    """The draws are keyed by the GLOBAL chain id (csrc/rng.cuh counters), so running all chains at once or as two
    shards (as two ranks would) gives the same states and the same trees."""
    S, C = 6, 10
    vg_np, vg_t = _target(S, 4)
    z0 = torch.as_tensor(np.random.default_rng(1).standard_normal((C, S)) * 0.3)
    run = lambda z, ids: nuts.nuts_run_(z, torch.full((len(ids),), 0.2, dtype=torch.float64),
                                        torch.zeros((len(ids), 4), dtype=torch.float64), lambda zz, bt: vg_t(zz),
                                        n_iter=4, fixed_beta_temp=1.0, seed=11, chain_ids=ids, max_tree_depth=5)
    ids = torch.arange(40, 40 + C, dtype=torch.int64)
    za = z0.clone(); oa = run(za, ids)
    zb1, zb2 = z0[:4].clone(), z0[4:].clone()
    ob1, ob2 = run(zb1, ids[:4]), run(zb2, ids[4:])
    assert torch.equal(za, torch.cat([zb1, zb2]))
    assert torch.equal(oa["n_leapfrog"], torch.cat([ob1["n_leapfrog"], ob2["n_leapfrog"]], dim=1))
