"""The C restatement (oracle/magi_oracle_c.c) against the numpy oracle: log-posterior + gradient for every compiled-in
system, leapfrog trajectories, whole HMC and NUTS chains draw for draw."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import magi_oracle as mo
from tests.helpers import random_state, relerr, seir_vignette_constants, synth_constants


@pytest.mark.parametrize("model", list(mo.MODELS))
@pytest.mark.parametrize("band", [None, 6])
def test_logpost_grad(model, band):
    rng = np.random.default_rng(11)
    c = synth_constants(model, 1, N=9, disc=1, band=6)
    O = co.COracle(c, model, band=band)
    X, s, tau = random_state(c, model, rng, 3)
    for r in range(3):
        lp, gX, gs, gt = mo.log_posterior_and_grad_analytic(X[r], s[r], tau[r], 0.7, c, model)
        lpc, gc = O.logpost_grad(mo.pack_state(X[r], s[r], tau[r]), 0.7)
        assert abs(lp - lpc) <= 1e-12 * abs(lp)
        assert relerr(gc, mo.pack_state(gX, gs, gt)) <= 1e-12
    lpb, Gb = O.logpost_grad_batch(np.stack([mo.pack_state(X[r], s[r], tau[r]) for r in range(3)]), 0.7)
    assert abs(lpb[2] - lpc) <= 1e-15 * abs(lpc) and np.array_equal(Gb[2], gc)


def test_vignette_shape_against_autograd():
    c, _, _ = seir_vignette_constants()
    rng = np.random.default_rng(5)
    X, s, tau = random_state(c, "seir3", rng, 1)
    lp, gX, gs, gt = mo.log_posterior_and_grad_autograd(X[0], s[0], tau[0], 0.37, c)
    lpc, gc = co.COracle(c, "seir3", band=80).logpost_grad(mo.pack_state(X[0], s[0], tau[0]), 0.37)
    assert abs(lp - lpc) <= 1e-11 * abs(lp)
    assert relerr(gc, mo.pack_state(gX, gs, gt)) <= 1e-11


def test_leapfrog_hmc_nuts_draw_for_draw():
    model = "seir3"
    rng = np.random.default_rng(2)
    c = synth_constants(model, 2, N=9, disc=1, band=6)
    n, D, P = c.n, 3, 3
    O = co.COracle(c, model, band=6)
    X, s, tau = random_state(c, model, rng, 1)
    z0 = mo.pack_state(X[0], s[0], tau[0])

    def vg(zz):
        Xx, ss, tt = mo.unpack_state(zz, n, D, P)
        lp, gX, gs, gt = mo.log_posterior_and_grad_analytic(Xx, ss, tt, 0.6, c, model)
        return lp, mo.pack_state(gX, gs, gt)

    p0 = mo.rng_normals(7, 1, 0, len(z0))
    z1, p1, _, _, traj = mo.leapfrog(z0, p0, 2e-3, 6, vg)
    zc, pc, trajc = O.leapfrog(z0, p0, 2e-3, 6, 0.6)
    assert relerr(trajc, np.array(traj)) <= 1e-12 and relerr(pc, p1) <= 1e-11

    zs, acc, eps, _ = mo.hmc_chain(c, model, z0, 10, 5, 0.01, seed=5, chain_id=3, num_adaptation_steps=8)
    r = O.hmc_chain(z0, 10, 5, 0.01, seed=5, chain_id=3, num_adaptation_steps=8, store_z=True)
    assert relerr(r["z"], zs) <= 1e-10 and np.allclose(r["accept"], acc, atol=1e-9) and np.allclose(r["step_size"], eps, rtol=1e-10)

    zs, acc, eps, nl = mo.nuts_chain(c, model, z0, 12, 0.01, seed=5, chain_id=3, num_adaptation_steps=8, max_tree_depth=6)
    r = O.nuts_chain(z0, 12, 0.01, seed=5, chain_id=3, num_adaptation_steps=8, max_tree_depth=6, store_z=True)
    assert np.array_equal(r["leapfrogs"], nl)
    assert relerr(r["z"], zs) <= 1e-9 and np.allclose(r["accept"], acc, atol=1e-9) and np.allclose(r["step_size"], eps, rtol=1e-9)
    assert np.allclose(r["tail"], zs[:, n * D:], rtol=1e-9, atol=1e-12)


def test_cached_lp_mode_differs_only_through_the_temperature_change():
    """cached_lp = 1 reuses the previous transition's log-posterior / gradient (TFP kernel results): identical chains at
    a fixed temperature, different ones under the annealing schedule."""
    model = "seir3"
    c = synth_constants(model, 4, N=9, disc=1, band=6)
    O = co.COracle(c, model, band=6)
    X, s, tau = random_state(c, model, np.random.default_rng(3), 1)
    z0 = mo.pack_state(X[0], s[0], tau[0])
    a = O.nuts_chain(z0, 8, 0.01, seed=1, fixed_beta_temp=0.5, max_tree_depth=5, cached_lp=False, store_z=True)
    b = O.nuts_chain(z0, 8, 0.01, seed=1, fixed_beta_temp=0.5, max_tree_depth=5, cached_lp=True, store_z=True)
    assert relerr(b["z"], a["z"]) <= 1e-12
    a = O.nuts_chain(z0, 8, 0.01, seed=1, max_tree_depth=5, cached_lp=False, store_z=True)
    b = O.nuts_chain(z0, 8, 0.01, seed=1, max_tree_depth=5, cached_lp=True, store_z=True)
    assert relerr(b["z"], a["z"]) > 1e-9
