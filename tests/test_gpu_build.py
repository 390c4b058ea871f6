"""Parity of the CUDA covariance build and Cholesky-based factorisation with the reference's
`_build_matrices` (magi_v2.py:774-823) and the SVD pseudo-inverses at :126-128.

Tolerances (hard part 1): the Matern blocks C, C', C'' are compared entry-wise
(1e-12 of the block's scale); m, K, C^-1, K^-1 inherit a forward error ~ eps * cond(C) from the
reference's own pinv route, so they are compared with tolerance c * eps * cond(C)."""
import numpy as np
import pytest

from oracle import magi_oracle as mo
from tests.helpers import load_golden, relerr

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def _T(a, device):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)


def _cov(I, phi1, phi2, nu, device, uniform=False):
    import torch
    from magi_v2_b200 import ops
    C, Cp, Cpp = ops.cov_build(_T(I, device), _T(phi1, device), _T(phi2, device), nu, uniform)
    torch.cuda.synchronize()
    return C.cpu().numpy(), Cp.cpu().numpy(), Cpp.cpu().numpy()


def _assert_blocks_match_reference(C, Cp, Cpp, I, p1, p2, v, Kap, pK, Kpp):
    """C: flat 1e-13.  C', C'': within the reference's own rounding-error bound (its formulas cancel,
    see oracle.matern_blocks_roundoff_scale) -- the CUDA path uses cancellation-free forms."""
    M_pK, M_Kpp = mo.matern_blocks_roundoff_scale(I, p1, p2, v)
    assert relerr(C, Kap) <= 1e-13
    assert np.all(np.abs(Cp - pK) <= 32 * EPS * M_pK + 1e-13 * np.abs(pK).max())
    assert np.all(np.abs(Cpp - Kpp) <= 32 * EPS * M_Kpp + 1e-13 * np.abs(Kpp).max())
    assert np.array_equal(Cp, -Cp.T) and np.array_equal(C, C.T) and np.array_equal(Cpp, Cpp.T)


@pytest.mark.parametrize("tag", ["appB", "n21", "n41", "n33nu25", "n17ragged"])
def test_matern_blocks_match_reference_golden(tag, cuda_device):
    g = load_golden("build_kat.npz")
    I = g[f"{tag}_I"]
    p1, p2, v = g[f"{tag}_hp"]
    C, Cp, Cpp = _cov(I, [[p1]], [[p2]], float(v), cuda_device)
    _assert_blocks_match_reference(C[0, 0], Cp[0, 0], Cpp[0, 0], I, p1, p2, float(v), g[f"{tag}_C"],
                                   g[f"{tag}_pK"], g[f"{tag}_Kpp"])    # _C is genuine reference output


@pytest.mark.parametrize("n,phi2", [(161, 0.375), (161, 0.109), (321, 0.23)])
def test_matern_blocks_full_size_probes(n, phi2, cuda_device):
    g = load_golden("build_kat.npz")
    tag = f"n{n}_phi2_{phi2}"
    I = np.linspace(0, 4, n)
    M_pK, M_Kpp = mo.matern_blocks_roundoff_scale(I, 0.0085, phi2, 2.01)
    for uniform in (False, True):
        C, Cp, Cpp = _cov(I, [[0.0085]], [[phi2]], 2.01, cuda_device, uniform)
        for A, nm, M in ((C, "C", None), (Cp, "pK", M_pK), (Cpp, "Kpp", M_Kpp)):
            A = A[0, 0]
            scale = np.abs(g[f"{tag}_{nm}_row0"]).max()
            slack = (lambda r: 1e-13 * scale) if M is None else (lambda r: 32 * EPS * r + 1e-13 * scale)
            assert np.all(np.abs(A[0] - g[f"{tag}_{nm}_row0"]) <= slack(M[0] if M is not None else 0))
            assert np.all(np.abs(A[n // 2] - g[f"{tag}_{nm}_rowmid"]) <= slack(M[n // 2] if M is not None else 0))
            assert np.all(np.abs(np.diag(A) - g[f"{tag}_{nm}_diag"]) <= 1e-13 * scale)
        # Toeplitz consistency that the reference itself lacks: equal lags give equal entries
        if uniform:
            Q = Cpp[0, 0]
            assert np.array_equal(Q[1:, :-1].diagonal(), np.full(n - 1, Q[1, 0]))


def test_batched_build_many_datasets(cuda_device):
    rng = np.random.default_rng(0)
    B, D, n = 5, 3, 37
    I = np.sort(rng.uniform(0, 3, (B, n)), axis=1)
    phi1 = rng.uniform(0.005, 0.05, (B, D)); phi2 = rng.uniform(0.1, 0.5, (B, D))
    C, Cp, Cpp = _cov(I, phi1, phi2, 2.01, cuda_device)
    for b in range(B):
        for d in range(D):
            Kap, pK, Kpp = mo.matern_blocks(I[b], phi1[b, d], phi2[b, d], 2.01)
            _assert_blocks_match_reference(C[b, d], Cp[b, d], Cpp[b, d], I[b], phi1[b, d], phi2[b, d], 2.01,
                                           Kap, pK, Kpp)


def test_matern_blocks_against_mpmath_truth(cuda_device):
    """Independent 40-digit ground truth at ragged lags, including very small ones."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    I = np.array([0.0, 1e-3, 0.0135, 0.05, 0.3, 1.1, 2.9, 4.0])
    p1, p2, nu = 0.02, 0.23, 2.01
    C, Cp, Cpp = _cov(I, [[p1]], [[p2]], nu, cuda_device)
    a = mp.sqrt(2 * mp.mpf(nu)) / mp.mpf(p2)
    c = mp.mpf(p1) * 2 ** (1 - mp.mpf(nu)) / mp.gamma(mp.mpf(nu))
    for i in range(len(I)):
        for j in range(len(I)):
            if i == j:
                continue
            l = abs(mp.mpf(float(I[i])) - mp.mpf(float(I[j])))
            u = a * l
            t0 = c * u ** nu * mp.besselk(nu, u)
            t1 = -c * a * u ** nu * mp.besselk(nu - 1, u) * (1 if I[i] > I[j] else -1)
            t2 = -c * a * a * u ** (nu - 1) * (u * mp.besselk(nu - 2, u) - mp.besselk(nu - 1, u))
            for got, want in ((C[0, 0, i, j], t0), (Cp[0, 0, i, j], t1), (Cpp[0, 0, i, j], t2)):
                assert abs(got - float(want)) <= 5e-13 * abs(float(want)) + 1e-300


@pytest.mark.parametrize("n,phi2,band", [(21, 0.375, None), (41, 0.23, 10), (161, 0.109, 80), (161, 0.375, 80),
                                          (130, 0.3, None)])
def test_factor_derive_matches_reference_route(n, phi2, band, cuda_device):
    """Both routes (reference: SVD pinv, here: Cholesky) are measured against an extended-precision
    solve of the same double-precision blocks; the CUDA result must be as close to it as the reference
    is (x5), and the two must agree within a cond(C)-scaled bound."""
    import torch
    from magi_v2_b200 import ops
    from tests.helpers import matern_truth_longdouble
    I = np.linspace(0, 4.0 * (n - 1) / 160.0 if n != 21 else 1.0, n)
    phi1 = 0.0085
    C_ref, m_ref, K_ref = mo.build_matrices(I, phi1, phi2, 2.01)          # reference route (pinv)
    Cinv_ref, Kinv_ref = mo.tf_pinv(C_ref), mo.tf_pinv(K_ref)
    Cinv_t, m_t, K_t, Kinv_t = (np.asarray(a, dtype=np.float64) for a in matern_truth_longdouble(I, phi1, phi2, 2.01))
    cond = np.linalg.cond(C_ref)
    Kap, pK, Kpp = mo.matern_blocks(I, phi1, phi2, 2.01)
    Cinv, m, Kinv, K, info = ops.factor_derive(_T(Kap[None], cuda_device), _T(pK[None], cuda_device),
                                               _T(Kpp[None], cuda_device), -1 if band is None else band, 0.0)
    torch.cuda.synchronize()
    assert int(info[0]) == 0
    Cinv, m, Kinv, K = (a.cpu().numpy()[0] for a in (Cinv, m, Kinv, K))
    bp = lambda A: mo.band_part(A, band)
    for mine, ref, truth, nm in ((K, K_ref, K_t, "K"), (m, bp(m_ref), bp(m_t), "m"), (Cinv, bp(Cinv_ref), bp(Cinv_t), "Cinv"),
                                 (Kinv, bp(Kinv_ref), bp(Kinv_t), "Kinv")):
        e_ref, e_mine = relerr(ref, truth), relerr(mine, truth)
        assert e_mine <= 5 * e_ref + 1e-12, (nm, e_mine, e_ref)
        assert relerr(mine, ref) <= 2e3 * EPS * cond * (np.linalg.cond(K_ref) if nm == "Kinv" else 1.0), nm
    if band is None:
        assert np.abs(Cinv @ C_ref - np.eye(n)).max() <= 50 * EPS * cond
        assert np.abs(Kinv @ K - np.eye(n)).max() <= 1e-10


def test_factor_reports_non_positive_definite(cuda_device):
    import torch
    from magi_v2_b200 import ops
    n = 9
    A = np.eye(n); A[4, 4] = -1.0
    Z = np.zeros((1, n, n))
    *_, info = ops.factor_derive(_T(A[None], cuda_device), _T(Z, cuda_device), _T(np.eye(n)[None], cuda_device), -1, 0.0)
    torch.cuda.synchronize()
    assert int(info[0]) == 5


def test_end_to_end_build_feeds_the_log_posterior(cuda_device):
    """cov_build -> factor_derive -> pack -> logpost_grad entirely on the device, checked against the
    oracle evaluated on the DEVICE-built matrices (isolates the evaluation from the pinv-vs-Cholesky
    conditioning difference)."""
    import torch
    from magi_v2_b200 import ops
    from tests.helpers import random_state, synth_constants
    model = "seir4"
    c0 = synth_constants(model, seed=77, N=21, band=12)
    rng = np.random.default_rng(4)
    D = 4
    phi1 = rng.uniform(0.005, 0.05, (1, D)); phi2 = rng.uniform(0.2, 0.6, (1, D))
    C, Cp, Cpp = ops.cov_build(_T(c0.I.ravel(), cuda_device), _T(phi1, cuda_device), _T(phi2, cuda_device), 2.01, True)
    Cinv, m, Kinv, K, info = ops.factor_derive(C, Cp, Cpp, 12, 0.0)
    assert int(info.abs().max()) == 0
    c = mo.PosteriorConstants(I=c0.I, mu_ds=c0.mu_ds, C_d_invs=Cinv[0].cpu().numpy(), m_ds=m[0].cpu().numpy(),
                              K_d_invs=Kinv[0].cpu().numpy(), N_ds=c0.N_ds, not_nan_idxs=c0.not_nan_idxs,
                              not_nan_cols=c0.not_nan_cols, y_tau_ds_observed=c0.y_tau_ds_observed, beta=c0.beta,
                              sigma_sqs_LB=c0.sigma_sqs_LB, f_vec=c0.f_vec)
    from tests.helpers import device_problem
    prob = device_problem([c], model, cuda_device)
    X, s, tau = random_state(c, model, rng, 4)
    lp, gX, gs, gt = prob.logpost_grad(_T(X[None], cuda_device), _T(s[None], cuda_device), _T(tau[None], cuda_device),
                                       _T(np.full((1, 4), 0.8), cuda_device))
    torch.cuda.synchronize()
    for r in range(4):
        o = mo.log_posterior_and_grad_autograd(X[r], s[r], tau[r], 0.8, c)
        assert abs(float(lp[0, r]) - o[0]) <= 1e-9 * abs(o[0])
        assert relerr(gX[0, r].cpu().numpy(), o[1]) <= 1e-9


@pytest.mark.parametrize("n,nmat", [(41, 7), (81, 5), (130, 3), (161, 4)])
def test_spd_inverse_matches_lapack(n, nmat, cuda_device):
    """`magi_b200_spd_inverse` (the hyper-parameter fit's S^-1 and log det S, magi_v2.py:594-597) against LAPACK on
    the GP covariances it is used for: Matern block + noise."""
    import torch
    from magi_v2_b200 import ops
    rng = np.random.default_rng(n)
    I = np.linspace(0.0, 4.0, n)
    A = np.stack([mo.matern_blocks(I, rng.uniform(0.01, 2.0), rng.uniform(0.2, 1.0), 2.01)[0] +
                  rng.uniform(1e-4, 1e-2) * np.eye(n) for _ in range(nmat)])
    Ainv, logdet, info = ops.spd_inverse(_T(A, cuda_device))
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0
    for k in range(nmat):
        ref = np.linalg.inv(A[k])
        assert relerr(Ainv[k].cpu().numpy(), ref) <= 1e-13 * np.linalg.cond(A[k])
        assert abs(float(logdet[k]) - np.linalg.slogdet(A[k])[1]) <= 1e-9 * n
    bad = A.copy()
    bad[1, 3, 3] = -1.0
    _, _, info = ops.spd_inverse(_T(bad, cuda_device))
    assert int(info[1]) == 4 and int(info[0]) == 0
