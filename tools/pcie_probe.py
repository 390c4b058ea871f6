"""What the host<->device link of this box delivers for the e2e leg of bench.py: pinned H2D alone, D2H alone, both at once
(separate streams), for the chunk size the pipeline uses (21 MB) and the whole step (171 MB).
    python tools/pcie_probe.py"""
import time

import torch

dev = torch.device("cuda:0")
for mb in (21, 171):
    n = mb * 1024 * 1024 // 8
    h_in = torch.empty(n, dtype=torch.float64).pin_memory()
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    d_in = torch.empty(n, dtype=torch.float64, device=dev)
    d_out = torch.empty(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h, reps=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return n * 8 * reps / (time.perf_counter() - t0) / 1e9

    run(True, True, 2)
    print(f"{mb:4d} MB  H2D alone {run(True, False):6.1f} GB/s   D2H alone {run(False, True):6.1f} GB/s   "
          f"both at once {run(True, True):6.1f} GB/s each way")
