"""BASELINE config 5 alone (Lorenz-96, n = 1281, D = 10, 2 datasets x 64 chains, dense matrices) through the wide path:
for ncu captures of wide_pass1/2/3.   python tools/wide_l96.py [reps]"""
import sys

import numpy as np
import torch

from magi_v2_b200 import ops

dev = torch.device("cuda:0")
T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
rng = np.random.default_rng(9)
n, D, B, R = 1281, 10, 2, 64
I = np.linspace(0, 4, n)
phi1, phi2 = rng.uniform(0.5, 2.0, (B, D)), rng.uniform(0.15, 0.3, (B, D))
C, Cp, Cpp = ops.cov_build(T(I), T(phi1), T(phi2), 2.01, True)
Cinv, m, Kinv, _, info = ops.factor_derive(C, Cp, Cpp, -1, 0.0)
assert int(info.abs().max()) == 0
packed = ops.pack_matrices(Cinv, m, Kinv)
del C, Cp, Cpp, Cinv, m, Kinv
mask = np.zeros((B, n, D), dtype=np.uint8); mask[:, ::16] = 1
y = rng.normal(2.0, 3.0, (B, n, D)) * mask
prob = ops.PosteriorProblem("lorenz96", packed, mu=T(np.full((B, D), 2.0)), y=T(y), mask=T(mask, torch.uint8),
                            N_ds=T(np.full((B, D), 81.0)), beta=T(np.full(B, D * n / (81.0 * D))),
                            LB=T(np.full((B, D), 1e-4)), n=n, band=None)
X = T(rng.normal(2.0, 3.0, (B, R, n, D))); s = T(rng.normal(-1, 0.5, (B, R, D))); tau = T(rng.normal(2.0, 0.2, (B, R, 1)))
bt = T(np.full((B, R), 1.0))
out = prob.logpost_grad_out(R)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for _ in range(3):
    prob.logpost_grad(X, s, tau, bt, out=out, path="wide")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    prob.logpost_grad(X, s, tau, bt, out=out, path="wide")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"L96 n=1281 dense B=2 R=64 wide: {ms:.3f} ms per evaluation sweep, {B * R / ms * 1e3:.0f} evals/s, "
      f"{8.0 * D * n * n * B * R / ms / 1e9:.2f} TFLOP/s")
