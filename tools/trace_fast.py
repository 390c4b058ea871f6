"""Timeline of CTA 0 of the fast log-posterior kernel, from a -DMAGI_TRACE build of the library:
    tools/build_variant.sh tr "-DMAGI_DEV_SEIR4_ONLY -DMAGI_TRACE"
    MAGI_B200_LIB=$PWD/variants/libmagi_tr.so python tools/trace_fast.py [B]
Tags: 10 eval begin | 11..17 after the barriers of fast_eval | 20 gradient stored | 21 next item loaded | per chunk: 2 wait begin, 3 data there, 4 refill issued,
5 contracted.  Prints, per warp, where the cycles of one evaluation go."""
import ctypes as C
import sys

import numpy as np
import torch

from magi_v2_b200 import _lib, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 6
dev = torch.device("cuda:0")
prob, info, state, data = synth.sweep_problem(B, 8, dev, seed0=0, model="seir4", bandsize=80)
X, s, tau = (torch.as_tensor(state[k], dtype=torch.float64, device=dev) for k in ("X", "sig_pre", "th_pre"))
bt = torch.full((B, 8), 0.37, dtype=torch.float64, device=dev)
L = _lib.lib()
fn = L.magi_b200_debug_trace
fn.restype = C.c_int
CAP, NW = 8192, 21
ev = np.zeros(NW * CAP, dtype=np.uint64)
cnt = np.zeros(NW, dtype=np.int32)
for _ in range(2):
    prob.logpost_grad(X, s, tau, bt)
    torch.cuda.synchronize()
    fn(ev.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
ev = ev.reshape(NW, CAP)
for w in (0, 5, 10, 15, 20):
    e = ev[w, :cnt[w]]
    t, tag = (e >> np.uint64(8)).astype(np.int64), (e & np.uint64(255)).astype(int)
    starts = np.flatnonzero(tag == 10)
    if len(starts) < 4:
        continue
    a, b = starts[2], starts[3]          # the third evaluation of this CTA
    tt, gg = t[a:b + 1] - t[a], tag[a:b + 1]
    print(f"warp {w}: evaluation = {tt[-1]} cycles, {int((gg == 2).sum())} chunks")
    wait = compute = issue = 0
    for i in range(1, len(gg)):
        d = tt[i] - tt[i - 1]
        if gg[i] == 3: wait += d
        elif gg[i] == 4: issue += d
        elif gg[i] == 5: compute += d
    print(f"   in mbarrier waits {wait}, refill issue {issue}, contraction {compute}, rest {tt[-1] - wait - issue - compute}")
    marks = [(int(tt[i]), int(gg[i])) for i in range(len(gg)) if gg[i] >= 10]
    print("   barriers:", " ".join(f"{g}@{c}" for c, g in marks))
    if w == 10:
        line = []
        for i in range(1, len(gg)):
            if gg[i] in (2, 3, 4, 5):
                line.append(f"{gg[i]}:{tt[i] - tt[i - 1]}")
        print("   chunk events (tag:delta):", " ".join(line[:160]))

# idle cycles of every warp before each barrier of the evaluation is released (release time - the warp's last event)
print("idle before barrier release (rows: warp, columns: barriers of the third evaluation in order)")
for w in range(NW):
    e = ev[w, :cnt[w]]
    t, tag = (e >> np.uint64(8)).astype(np.int64), (e & np.uint64(255)).astype(int)
    starts = np.flatnonzero(tag == 10)
    if len(starts) < 4:
        continue
    a, b = starts[2], starts[3]
    idle = [int(t[i] - t[i - 1]) for i in range(a + 1, b + 1) if 11 <= tag[i] <= 17]
    print(f"  warp {w:2d}: " + " ".join(f"{v:6d}" for v in idle))
