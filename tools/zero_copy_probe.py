"""Does the evaluation kernel run faster end to end when it reads the chain states straight from pinned HOST memory and
writes the gradients straight back (zero copy: no cudaMemcpy, no chunk pipeline) than through HostPipeline?
    python tools/zero_copy_probe.py [B]"""
import ctypes as C
import sys
import time

import numpy as np
import torch

from magi_v2_b200 import _lib, synth
from magi_v2_b200.ops import _ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = 8
dev = torch.device("cuda:0")
prob, info, state, data = synth.sweep_problem(B, R, dev, seed0=0, model="seir4", bandsize=80)
pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).pin_memory()
hX, hs, ht, hbt = pin(state["X"]), pin(state["sig_pre"]), pin(state["th_pre"]), pin(np.full((B, R), 0.37))
n, D, P = prob.n, prob.D, prob.P
olp, ogX, ogs, ogt = (torch.empty(s, dtype=torch.float64).pin_memory() for s in ((B, R), (B, R, n, D), (B, R, D), (B, R, P)))
fn, ws, nb = prob.eval_call(R, "cta")
pb = prob.struct(R)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def zero_copy():
    _lib.check(fn(C.byref(pb), _ptr(hX), _ptr(hs), _ptr(ht), _ptr(hbt), _ptr(olp), _ptr(ogX), _ptr(ogs), _ptr(ogt),
                  _ptr(ws), nb, st), "logpost_grad")


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


dX, ds, dt, dbt = (a.to(dev) for a in (hX, hs, ht, hbt))
ref = [t.cpu() for t in prob.logpost_grad(dX, ds, dt, dbt)]
zero_copy()
torch.cuda.synchronize()
err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip((olp, ogX, ogs, ogt), ref))
ms_zc = timed(zero_copy)
hp = prob.host_pipeline(R)
hp.fill(hX, hs, ht, hbt)
ms_hp = timed(hp.run)
ms_dev = timed(lambda: prob.logpost_grad(dX, ds, dt, dbt))
print(f"B={B}: device-resident {ms_dev:.3f} ms | HostPipeline {ms_hp:.3f} ms = {B * R / ms_hp / 1e3:.2f} M evals/s | "
      f"zero copy {ms_zc:.3f} ms = {B * R / ms_zc / 1e3:.2f} M evals/s (max rel diff vs device path {err:.1e})")
