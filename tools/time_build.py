"""Times the set-up kernels (cov_build, factor_derive, pack) on the device: CUDA events, 3 warm-ups.
   python tools/time_build.py [n] [nmat_datasets] [D]"""
import sys
import numpy as np
import torch
from magi_v2_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 161
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
D = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
I = T(np.linspace(0, 4, n))
p1, p2 = T(rng.uniform(0.005, 0.05, (B, D))), T(rng.uniform(0.1, 0.4, (B, D)))


def timed(fn, reps=5):
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


C, Cp, Cpp = ops.cov_build(I, p1, p2, 2.01, True)
ms_cov = timed(lambda: ops.cov_build(I, p1, p2, 2.01, True))
ms_cov_full = timed(lambda: ops.cov_build(I, p1, p2, 2.01, False))
ms_fac = timed(lambda: ops.factor_derive(C, Cp, Cpp, n // 2, 0.0))
out = ops.factor_derive(C, Cp, Cpp, n // 2, 0.0)
assert int(out[4].abs().max()) == 0
nmat = B * D
print(f"n={n} matrices={nmat}: cov_build uniform {ms_cov:.2f} ms, general {ms_cov_full:.2f} ms; "
      f"factor_derive {ms_fac:.2f} ms = {nmat * 6.0 * n ** 3 / (ms_fac * 1e-3) / 1e12:.2f} TFLOP/s (6 n^3 per matrix), "
      f"{nmat / (ms_fac * 1e-3):.0f} matrices/s")
