"""A/B of the two grid shapes of the evaluation (magi_b200_logpost_grad vs magi_b200_logpost_grad_wide) at larger
batches.  python tools/path_ab.py [sirw]"""
import sys

import numpy as np
import torch
from magi_v2_b200 import ops
dev = torch.device("cuda:0")
T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
def run(model, D, P, n, B, R, band):
    rng = np.random.default_rng(0)
    I = np.linspace(0, 4, n)
    phi1, phi2 = rng.uniform(0.01, 0.05, (B, D)), rng.uniform(0.15, 0.3, (B, D))
    npad = (n + 7) // 8 * 8
    packed = torch.empty(B * D * 3 * npad * npad, dtype=torch.float64, device=dev)
    per = D * 3 * npad * npad
    chunk = max(1, min(B, 2 ** 31 // (8 * D * n * n * 8)))
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        C, Cp, Cpp = ops.cov_build(T(I), T(phi1[b0:b1]), T(phi2[b0:b1]), 2.01, True)
        Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, -1 if band is None else band, 0.0)
        packed[b0 * per:b1 * per] = ops.pack_matrices(Cinv, m, Kinv)
        del C, Cp, Cpp, Cinv, m, Kinv
    mask = np.zeros((B, n, D), dtype=np.uint8); mask[:, ::(n - 1) // 80] = 1
    y = rng.normal(0.3, 0.1, (B, n, D)) * mask
    prob = ops.PosteriorProblem(model, packed, mu=T(np.full((B, D), 0.3)), y=T(y), mask=T(mask, torch.uint8),
                                N_ds=T(np.full((B, D), 81.0)), beta=T(np.full(B, D * n / (81.0 * D))),
                                LB=T(np.full((B, D), 1e-6)), n=n, band=band)
    X = T(rng.normal(0.3, 0.05, (B, R, n, D))); s = T(rng.normal(-6, 0.5, (B, R, D))); tau = T(rng.normal(0.5, 0.2, (B, R, P)))
    bt = T(np.full((B, R), 0.37))
    out = prob.logpost_grad_out(R)
    for path in ("cta", "wide"):
        for _ in range(3): prob.logpost_grad(X, s, tau, bt, out=out, path=path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): prob.logpost_grad(X, s, tau, bt, out=out, path=path)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{model} n={n} B={B} R={R} band={band} {path}: {ms:.3f} ms -> {B*R/ms*1e3:.3e} evals/s", flush=True)
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sirw":
        run("sirw", 4, 5, 321, 512, 8, None)
    else:
        run("sirw", 4, 5, 321, 512, 8, None)
        run("sirw", 4, 5, 321, 512, 8, 160)
        run("lorenz96", 10, 1, 1281, 8, 8, None)
        run("seir4", 4, 3, 161, 1024, 8, 80)
