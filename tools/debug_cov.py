import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import magi_oracle as mo
from magi_v2_b200 import ops
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
for n, p1, p2, T_ in ((21, 0.0085, 0.375, 1.0), (161, 0.0085, 0.375, 4.0)):
    I = np.linspace(0, T_, n)
    Kap, pK, Kpp = mo.matern_blocks(I, p1, p2, 2.01)
    C, Cp, Cpp = (a.cpu().numpy()[0, 0] for a in ops.cov_build(T(I), T([[p1]]), T([[p2]]), 2.01, False))
    for nm, A, Bm in (("C", C, Kap), ("Cp", Cp, pK), ("Cpp", Cpp, Kpp)):
        E = np.abs(A - Bm)
        i, j = np.unravel_index(E.argmax(), E.shape)
        print(n, nm, "max abs err", E.max(), "at", (i, j), "val", Bm[i, j], "mine", A[i, j], "scale", np.abs(Bm).max())
        print("   row0 errs:", E[0, :6])
