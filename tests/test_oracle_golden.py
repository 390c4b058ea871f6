"""Pin the oracle (oracle/magi_oracle.py) against the committed golden vectors, which were produced by
the GENUINE reference code (magi_v2.py) in the build container (oracle/make_golden.py), and -- when
/root/reference is present -- against the genuine code directly."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from oracle.ref_loader import reference_available
from tests.helpers import load_golden, relerr, seir_vignette_constants


@pytest.mark.parametrize("tag", ["appB", "n21", "n41", "n33nu25", "n17ragged"])
def test_build_matrices_restatement_matches_genuine_golden(tag):
    g = load_golden("build_kat.npz")
    p1, p2, v = g[f"{tag}_hp"]
    C, m, K = mo.build_matrices(g[f"{tag}_I"], p1, p2, v)
    assert np.array_equal(C, g[f"{tag}_C"])                      # elementwise: bit-exact
    # pinv / GEMM go through LAPACK/BLAS: allow for thread-count dependent summation order
    assert relerr(m, g[f"{tag}_m"]) <= 1e-9 and relerr(K, g[f"{tag}_K"]) <= 1e-9


def test_appendix_b_known_answer():
    """values printed by the genuinely executed reference."""
    C, m, K = mo.build_matrices(np.array([0, .025, .05, .075, .1]), 0.02, 0.23, 2.01)
    assert np.allclose(C[0], [0.02, 0.01977141381434577, 0.01913552684312188, 0.01818722654541616,
                              0.01702290374189537], rtol=1e-14)
    assert np.allclose(m[0], [-54.56023712813415, 76.6911004575541, -29.111890830634103, 8.962228479114154,
                              -1.8218661274944097], rtol=1e-9)
    assert np.allclose(np.diag(K), [0.03020075894921137, 0.01104144391102835, 0.00953313886470863,
                                    0.01104144391101081, 0.03020075894920771], rtol=1e-8)


def test_matern_blocks_against_mpmath():
    """Independent ground truth for kappa, kappa', kappa'' (SURVEY.md Appendix B)."""
    Kap, pK, Kpp = mo.matern_blocks(np.array([0.0, 0.025, 0.1]), 0.02, 0.23, 2.01)
    assert abs(Kap[0, 1] - 0.0197714138143457764) < 1e-16
    assert abs(pK[1, 0] - (-0.0178680141914813917)) < 1e-15      # d/ds at s>t equals kappa'(l)
    assert abs(Kpp[0, 1] - 0.655978164734165335) < 1e-13         # = -kappa''(l)
    assert abs(Kap[0, 2] - 0.0170229037418953702) < 1e-16
    assert abs(Kpp[0, 2] - 0.207690138391908756) < 1e-13


def test_discretize_and_interpolate_match_genuine_golden():
    g = load_golden("grid_kat.npz")
    for disc in (0, 1, 2):
        I, Xd = mo.discretize(g["ts"], g["X"], disc)
        assert np.array_equal(I, g[f"I_d{disc}"])
        assert np.array_equal(Xd, g[f"Xd_d{disc}"], equal_nan=True)
        assert np.array_equal(mo.linear_interpolate(Xd), g[f"Xi_d{disc}"], equal_nan=True)


@pytest.mark.parametrize("name", list(mo.MODELS))
def test_log_posterior_three_implementations_agree_with_golden(name):
    g = load_golden("logpost_kat.npz")
    model = mo.MODELS[name]
    mats = (g[f"{name}_Cinv"], g[f"{name}_m"], g[f"{name}_Kinv"])
    c = mo.make_constants(g[f"{name}_ts"], g[f"{name}_X_obs"], 1, g[f"{name}_phi1"], g[f"{name}_phi2"],
                          int(g[f"{name}_band"]), model.f_vec, matrices=mats)
    X, s, tau, bt = g[f"{name}_X"], g[f"{name}_s"], g[f"{name}_tau"], float(g[f"{name}_bt"])
    lp_np = mo.log_posterior(X, s, tau, bt, c)
    lp_ag, gX, gs, gt = mo.log_posterior_and_grad_autograd(X, s, tau, bt, c)
    lp_an, gX2, gs2, gt2 = mo.log_posterior_and_grad_analytic(X, s, tau, bt, c, name)
    ref = float(g[f"{name}_lp"])
    for v in (lp_np, lp_ag, lp_an):
        assert abs(v - ref) <= 1e-12 * abs(ref)
    for a, b, k in ((gX, gX2, "gX"), (gs, gs2, "gs"), (gt, gt2, "gt")):
        assert relerr(a, g[f"{name}_{k}"]) <= 1e-11
        assert relerr(b, g[f"{name}_{k}"]) <= 1e-11


def test_gradient_against_finite_differences():
    c, _, _ = seir_vignette_constants(0, "seir3")
    rng = np.random.default_rng(0)
    X = mo.linear_interpolate(np.where(c.dense_y_mask()[1] > 0, c.dense_y_mask()[0], np.nan))
    s, tau, bt = np.array([-5., -4., -6.]), np.array([6.0, 0.1, 1.6]), 0.37
    lp, gX, gs, gt = mo.log_posterior_and_grad_autograd(X, s, tau, bt, c)
    # SURVEY.md A.2 probe values on this configuration
    assert abs(c.beta - 1.98765) < 1e-5 and len(c.not_nan_idxs) == 243
    h = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd = (mo.log_posterior(X, s, tau + e, bt, c) - mo.log_posterior(X, s, tau - e, bt, c)) / (2 * h)
        assert abs(fd - gt[k]) <= 1e-5 * max(1.0, abs(gt[k]))
        fd = (mo.log_posterior(X, s + e, tau, bt, c) - mo.log_posterior(X, s - e, tau, bt, c)) / (2 * h)
        assert abs(fd - gs[k]) <= 1e-5 * max(1.0, abs(gs[k]))
    for (i, d) in ((0, 0), (80, 1), (160, 2)):
        E = np.zeros_like(X); E[i, d] = h
        fd = (mo.log_posterior(X + E, s, tau, bt, c) - mo.log_posterior(X - E, s, tau, bt, c)) / (2 * h)
        assert abs(fd - gX[i, d]) <= 1e-4 * max(1.0, abs(gX[i, d]))


def test_temperature_schedule_and_initial_state():
    """magi_v2.py:833-835 (SURVEY.md A.4 values) and :373-383."""
    assert abs(mo.logarithmic_temperature_schedule(0) - 1.4426950408889634) < 1e-15
    assert abs(mo.logarithmic_temperature_schedule(1) - 0.9102392266268373) < 1e-15
    assert abs(mo.logarithmic_temperature_schedule(1000) - 0.14471) < 1e-4
    assert mo.logarithmic_temperature_schedule(30000) == 0.1
    X, s0, t0 = mo.initial_state(np.ones((3, 2)), np.array([0.5, 1e-9]), np.array([2.0, -1.0]), np.array([0.1, 0.1]))
    assert abs(np.log1p(np.exp(s0[0])) + 0.1 - 0.5) < 1e-15 and s0[1] == -5.0
    assert abs(np.log1p(np.exp(t0[0])) - 2.0) < 1e-15 and t0[1] == -5.0


def test_philox_known_answer():
    """Random123 KAT: philox4x32-10, counter = key = 0 and all-ones."""
    z = mo.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    z = mo.philox4x32_10(f, f[:, :2])[0]
    assert [hex(int(v)) for v in z] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_rng_normals_moments():
    z = mo.rng_normals(7, 3, 11, 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert np.array_equal(z[:10], mo.rng_normals(7, 3, 11, 10))


def test_leapfrog_is_reversible_and_hmc_runs():
    c = None
    from tests.helpers import synth_constants
    c = synth_constants("seir3", seed=1)
    n, D, P = c.n, 3, 3
    rng = np.random.default_rng(0)
    from tests.helpers import random_state
    X, s, tau = random_state(c, "seir3", rng, 1)
    z0 = mo.pack_state(X[0], s[0], tau[0])

    def vg(z):
        a = mo.log_posterior_and_grad_analytic(*mo.unpack_state(z, n, D, P), 0.5, c, "seir3")
        return a[0], mo.pack_state(a[1], a[2], a[3])

    p0 = rng.standard_normal(z0.shape)
    z1, p1, *_ = mo.leapfrog(z0, p0, 1e-4, 5, vg)
    z2, p2, *_ = mo.leapfrog(z1, -p1, 1e-4, 5, vg)
    assert relerr(z2, z0) < 1e-10 and relerr(-p2, p0) < 1e-10
    zs, acc, eps, lps = mo.hmc_chain(c, "seir3", z0, 6, 3, 2e-4, seed=1, chain_id=0, num_adaptation_steps=4)
    assert zs.shape == (6, len(z0)) and np.all((acc >= 0) & (acc <= 1))


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_restatement_is_bit_identical_to_genuine_reference():
    from oracle.ref_loader import reference_object
    ref = reference_object()
    for n, p1, p2 in ((161, 0.0085, 0.375), (81, 0.034, 0.23), (33, 0.024, 0.109)):
        I = np.linspace(0, 4, n).reshape(-1, 1)
        a = ref._build_matrices(I, p1, p2, 2.01)
        b = mo.build_matrices(I, p1, p2, 2.01)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
