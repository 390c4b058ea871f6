"""Throughput of the hot path on the other BASELINE.json configs (parity cases, not bench lines):
config 3 (SIRW n = 321), config 5 (Lorenz-96 n = 1281), plus SEIR3 n = 161.  Synthetic constants; device-built
matrices.  Usage (GPU box): python tools/run_configs.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from magi_v2_b200 import ops

dev = torch.device("cuda:0")
T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)


def run(model, D, P, n, B, R, band, steps=5, L=8, jitter=0.0):
    rng = np.random.default_rng(0)
    I = np.linspace(0, 4, n)
    phi1, phi2 = rng.uniform(0.01, 0.05, (B, D)), rng.uniform(0.15, 0.3, (B, D))
    t0 = time.time()
    npad = (n + 7) // 8 * 8
    packed = torch.empty(B * D * 3 * npad * npad, dtype=torch.float64, device=dev)
    per = D * 3 * npad * npad
    chunk = max(1, min(B, 2 ** 31 // (8 * D * n * n * 8)))
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        C, Cp, Cpp = ops.cov_build(T(I), T(phi1[b0:b1]), T(phi2[b0:b1]), 2.01, True)
        Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, -1 if band is None else band, jitter)
        assert int(info.abs().max()) == 0, info
        packed[b0 * per:b1 * per] = ops.pack_matrices(Cinv, m, Kinv)
        del C, Cp, Cpp, Cinv, m, Kinv
    torch.cuda.synchronize()
    t_build = time.time() - t0
    mask = np.zeros((B, n, D), dtype=np.uint8); mask[:, ::(n - 1) // 80] = 1
    y = rng.normal(0.3, 0.1, (B, n, D)) * mask
    prob = ops.PosteriorProblem(model, packed, mu=T(np.full((B, D), 0.3)), y=T(y), mask=T(mask, torch.uint8),
                                N_ds=T(np.full((B, D), 81.0)), beta=T(np.full(B, D * n / (81.0 * D))),
                                LB=T(np.full((B, D), 1e-6)), n=n, band=band)
    X = T(rng.normal(0.3, 0.05, (B, R, n, D))); s = T(rng.normal(-6, 0.5, (B, R, D))); tau = T(rng.normal(0.5, 0.2, (B, R, P)))
    bt = T(np.full((B, R), 0.37))
    out = prob.logpost_grad_out(R)
    for _ in range(2):
        prob.logpost_grad(X, s, tau, bt, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        prob.logpost_grad(X, s, tau, bt, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert torch.isfinite(out[0]).all()
    eps = torch.full((B, R), 1e-5, dtype=torch.float64, device=dev)
    da = torch.zeros((B, R, 4), dtype=torch.float64, device=dev)
    prob.hmc_run_(X, s, tau, eps, da, n_iter=1, n_leapfrog=L, fixed_beta_temp=0.37)
    torch.cuda.synchronize()
    e0.record()
    o = prob.hmc_run_(X, s, tau, eps, da, n_iter=2, n_leapfrog=L, fixed_beta_temp=0.37)
    e1.record(); torch.cuda.synchronize()
    ms_h = e0.elapsed_time(e1) / 2
    bytes_eval = 24.0 * D * n * n / R + 16.0 * (n * D + D + P)
    print(f"{model:9s} n={n:5d} D={D:2d} B={B:5d} R={R:3d} band={band}: build {t_build:6.2f}s | logpost {ms:9.3f} ms "
          f"-> {B*R/ms*1e3:10.3e} evals/s ({bytes_eval*B*R/ms/1e6:7.1f} GB/s alg., "
          f"{8.0*D*n*n*B*R/ms/1e9:6.2f} TF/s) | HMC L={L}: {B*R/ms_h*1e3:10.3e} samples/s, accept {float(o['accept_prob'].mean()):.2f}",
          flush=True)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        run("seir3", 3, 3, 161, 4096, 8, 80)
        run("seir4", 4, 3, 161, 4096, 8, 80)
        run("seir4", 4, 3, 161, 20, 8, 80)
        run("sirw", 4, 5, 321, 512, 8, None)
    for jit in (0.0, 1e-12, 1e-10, 1e-8):   # (needed 1e-8 before K was formed as C'' - W^T W)
        try:
            run("lorenz96", 10, 1, 1281, 2, 64, None, steps=3, L=4, jitter=jit)
            run("lorenz96", 10, 1, 1281, 8, 8, None, steps=3, L=4, jitter=jit)
            print("jitter", jit, "ok")
            break
        except AssertionError as e:
            print("jitter", jit, "factorisation failed:", str(e)[:200])
