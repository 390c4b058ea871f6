"""Evaluations/s of the two grid shapes of the log-posterior + gradient (magi_b200_logpost_grad: one CTA per
(dataset, 8 chains); magi_b200_logpost_grad_wide: matrix rows spread over the grid) for few datasets.
   python tools/time_wide.py"""
import numpy as np
import torch

from magi_v2_b200 import ops, synth

dev = torch.device("cuda:0")
T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, prob, X, s, tau, bt):
    B, R = X.shape[:2]
    out = prob.logpost_grad_out(R)
    res = {}
    for path in ("cta", "wide"):
        ms = timed(lambda: prob.logpost_grad(X, s, tau, bt, out=out, path=path))
        res[path] = [o.clone() for o in out]
        print(f"{name:34s} B={B:4d} R={R:3d} {path:5s}: {ms * 1e3:9.1f} us per launch, {B * R / (ms * 1e-3):12.0f} evals/s")
    err = max(float(((a - b).abs().max() / b.abs().max())) for a, b in zip(res["cta"], res["wide"]))
    print(f"{'':34s} max rel. difference between the two paths: {err:.2e}")


for B in (1, 4, 20, 64):
    prob, info, state, _ = synth.sweep_problem(B, 8, dev, seed0=0, model="seir4", bandsize=80)
    report("SEIR4 n=161 band 80", prob, T(state["X"]), T(state["sig_pre"]), T(state["th_pre"]), T(np.full((B, 8), 0.37)))

# config 5 shape: Lorenz-96, n = 1281, D = 10
rng = np.random.default_rng(9)
n, D = 1281, 10
for B, R in ((1, 8), (2, 64)):
    I = np.linspace(0, 4, n)
    phi1, phi2 = rng.uniform(0.5, 2.0, (B, D)), rng.uniform(0.15, 0.3, (B, D))
    C, Cp, Cpp = ops.cov_build(T(I), T(phi1), T(phi2), 2.01, True)
    Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, 320, 0.0)
    assert int(info.abs().max()) == 0
    packed = ops.pack_matrices(Cinv, m, Kinv)
    del C, Cp, Cpp, Cinv, m, Kinv
    mask = np.zeros((B, n, D), dtype=np.uint8); mask[:, ::16] = 1
    y = rng.normal(2.0, 3.0, (B, n, D)) * mask
    prob = ops.PosteriorProblem("lorenz96", packed, mu=T(np.full((B, D), 2.0)), y=T(y), mask=T(mask, torch.uint8),
                                N_ds=T(np.full((B, D), 81.0)), beta=T(np.full(B, D * n / (81.0 * D))),
                                LB=T(np.full((B, D), 1e-4)), n=n, band=320)
    X = rng.normal(2.0, 3.0, (B, R, n, D)); s = rng.normal(-1, 0.5, (B, R, D)); tau = rng.normal(2.0, 0.2, (B, R, 1))
    report("Lorenz-96 n=1281 band 320", prob, T(X), T(s), T(tau), T(np.full((B, R), 1.0)))
