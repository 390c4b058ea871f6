#!/usr/bin/env python
"""Benchmark of the MAGI posterior-evaluation hot path on B200 (driver contract: see README/DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (restated)

Workload (BASELINE.json configs[3]): synthetic SEIR-shaped sweep,
B = 4096 datasets x R = 8 chains PER GPU (weak scaling: every rank owns its own 4096 datasets, no
data-path collective), n = 161 grid points, D = 4 components, P = 3 parameters, band 80.
A "step" is one log-posterior + analytic-gradient evaluation of all B*R chains (one launch of
`magi_b200_logpost_grad`); `value` = evaluations/s with inputs resident in HBM; `e2e` = the same
through the host-buffer entry point `PosteriorProblem.host_pipeline` (chain states in pinned HOST
blocks, one host->device copy, one launch and one device->host copy per dataset chunk inside the
timed region).  Further objects on the line:

  roofline       the dominant kernel against the measured HBM roof (MEASURED_PEAKS.json), the FP64 roof measured at run
                 time with the library's DMMA probe (include/magi_b200_probe.h), DRAM traffic from profiles/traffic.json
  parity_spot    after the timed loop: k random (dataset, chain) outputs of the measured launch against the oracle
                 (compiled C restatement, oracle/magi_oracle_c.c) on device-built matrices
  build          the set-up path: ms per 2048 matrices for cov_build and factor_derive, with the reference route
                 (magi_v2.py:774-823 + two SVD pseudo-inverses, :126-128) timed beside it on the host cores
  hmc / nuts     transitions/s of the fused fixed-length sampler and of the reference's sampler stack (device and e2e)
  allgather      (N > 1) the one collective, timed alone
  cpu_baseline   the reference's CPU path on the host cores (bounded sample; see run_reference)
  other_configs  (rank 0) BASELINE configs 2, 3, 5: evaluation throughput + a parity spot check each

The reference arm: TensorFlow-Probability is not installable in this image, so it times the restatement of
the reference's path.  Its `value` is the compiled analytic-gradient port (C, all chains of a dataset per call, one
thread per core) -- the arm an XLA-compiled TFP graph is closest to; the eager torch-autograd restatement the first
round quoted is reported beside it as `autograd_eager`.  The oracle is executed by this file only there, in the
`cpu_baseline` leg and as the post-hoc checker of `parity_spot`; the measured CUDA path never imports it."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "magi_logpost_grad_evals_per_s"
UNIT = "evals/s"
N_GRID, D, P = 161, 4, 3
S_STATE = N_GRID * D + D + P
ORACLE_PIN = ("log-posterior / sampler oracle restates a TFP graph that cannot run in this image: parity unpinned "
              "(DESIGN.md section 2); covariance build pinned to the genuine reference code")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--datasets", type=int, default=4096, help="datasets per GPU")
    ap.add_argument("--chains", type=int, default=8, help="chains per dataset")
    ap.add_argument("--leapfrog", type=int, default=16, help="leapfrog steps per HMC transition")
    ap.add_argument("--hmc-iters", type=int, default=4, help="HMC transitions per timed HMC step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hmc", action="store_true")
    ap.add_argument("--no-nuts", action="store_true")
    ap.add_argument("--no-build", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short evaluation-throughput lines of BASELINE configs 2, 3 and 5")
    ap.add_argument("--nuts-depth", type=int, default=4, help="max_tree_depth of the timed NUTS transitions")
    ap.add_argument("--cpu-calls-per-thread", type=int, default=1500,
                    help="reference arm: dataset evaluations (8 chains each) per thread and step")
    ap.add_argument("--cpu-autograd-evals", type=int, default=12, help="reference arm: eager-autograd evaluations per worker")
    ap.add_argument("--e2e-chunks", type=int, default=int(os.environ.get("MAGI_E2E_CHUNKS", "16")))
    ap.add_argument("--e2e-streams", type=int, default=int(os.environ.get("MAGI_E2E_STREAMS", "4")))
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "synthetic SEIR4 sweep (BASELINE configs[3]): datasets x chains of independent MAGI posteriors",
            "datasets_per_gpu": args.datasets, "chains_per_dataset": args.chains, "n_grid": N_GRID, "D": D, "P": P,
            "bandsize": 80, "sharding": f"datasets over {n_gpus} rank(s), no data-path collective",
            "l2": "inputs (packed matrices, %.1f GB per GPU) exceed the 126 MB L2; no flush needed"
                  % (args.datasets * D * 3 * 168 * 168 * 8 / 1e9)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# reference arm: the restated reference CPU path on all host cores
# ------------------------------------------------------------------------------------------------
def _sweep_constants(seed0):
    """One SEIR4 dataset of the sweep as the oracle's PosteriorConstants; matrices by the reference route
    (scipy kvp + SVD pseudo-inverses, magi_v2.py:774-823, :126-128)."""
    import numpy as np
    from magi_v2_b200 import synth
    from oracle import magi_oracle as mo
    data = synth.seir_sweep(1, seed0=seed0, model="seir4")
    rng = np.random.default_rng(seed0)
    phi1, phi2 = rng.uniform(0.005, 0.05, D), rng.uniform(0.1, 0.4, D)
    c = mo.make_constants(data["ts_obs"], data["X_obs"][0], 1, phi1, phi2, 80, mo.f_seir4)
    y, mask = c.dense_y_mask()
    Xhat = mo.linear_interpolate(np.where(mask > 0, y, np.nan))
    return c, Xhat


_W = {}


def _autograd_worker_init(seed0):
    import torch
    torch.set_num_threads(1)
    try:
        from threadpoolctl import threadpool_limits
        _W["blas_limit"] = threadpool_limits(limits=1)       # one BLAS thread per worker process
    except ImportError:
        pass
    import numpy as np
    from oracle import magi_oracle as mo
    c, Xhat = _sweep_constants(seed0)
    _W.update(mo=mo, c=c, Xhat=Xhat, rng=np.random.default_rng(seed0))
    return True


def _autograd_worker_eval(n_evals):
    import numpy as np
    mo, c, Xhat, rng = _W["mo"], _W["c"], _W["Xhat"], _W["rng"]
    acc = 0.0
    for _ in range(n_evals):
        X = Xhat + 0.01 * rng.standard_normal(Xhat.shape)
        s = rng.normal(-4, 1, D)
        tau = np.log(np.expm1(np.array([6.0, 0.6, 1.8]) * np.exp(rng.uniform(-0.1, 0.1, 3))))
        lp, gX, gs, gt = mo.log_posterior_and_grad_autograd(X, s, tau, 0.37, c)
        acc += lp
    return acc


def autograd_eager_rate(workers, per, steps):
    """Op-for-op torch-CPU restatement of magi_v2.py:308-348 + autograd, one chain per call (as the reference's TFP loop
    calls its target), one process per core."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, initializer=_autograd_worker_init, initargs=(1000 + w,)) for w in range(workers)]
    run = lambda: [r.get() for r in [p.apply_async(_autograd_worker_eval, (per,)) for p in pools]]
    run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    for p in pools:
        p.close()
    return steps * workers * per / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import c_oracle as co
    from oracle import magi_oracle as mo
    cores = max(1, host_cores())
    R = args.chains
    # a few distinct datasets (the reference route takes ~0.7 s per dataset), one private copy per thread
    base = [_sweep_constants(1000 + w) for w in range(min(cores, 4))]
    rng = np.random.default_rng(0)
    work = []
    for w in range(cores):
        c, Xhat = base[w % len(base)]
        O = co.COracle(c, "seir4", band=80)
        Z = np.stack([mo.pack_state(Xhat + 0.01 * rng.standard_normal(Xhat.shape), rng.normal(-4, 1, D),
                                    np.log(np.expm1(np.array([6.0, 0.6, 1.8]) * np.exp(rng.uniform(-0.1, 0.1, 3)))))
                      for _ in range(R)])
        work.append((O, Z))
    per = args.cpu_calls_per_thread

    def thread_job(w):
        O, Z = work[w]
        acc = 0.0
        for _ in range(per):
            lp, G = O.logpost_grad_batch(Z, 0.37)            # ctypes releases the GIL: threads run in parallel
            acc += float(lp[0])
        return acc

    ex = ThreadPoolExecutor(cores)
    step = lambda: sum(ex.map(thread_job, range(cores)))
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ex.shutdown()
    value = args.steps * cores * per * R / dt
    eager = autograd_eager_rate(cores, args.cpu_autograd_evals, 2)
    sample = (f"{cores} threads x {per} dataset calls x {R} chains per step x {args.steps} steps; each call = the C "
              "restatement of unnormalized_log_prob (magi_v2.py:308-348) + analytic gradient (SURVEY.md A.3) for the "
              f"{R} chains of one SEIR4 dataset (n=161, D=4, band 80; matrices by the reference's SVD route), gcc -O3 "
              "AVX2, matrices cache-resident (a 4096-dataset sweep would stream them from DRAM)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "autograd_eager": {"value": eager, "unit": UNIT, "cores": cores,
                                                "sample": "torch CPU FP64 op-for-op restatement + autograd, one chain "
                                                          "per call, one process per core (round-1 baseline)"}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "oracle_pin": ORACLE_PIN,
            "note": "TensorFlow-Probability is not installable in this image; kind=port: compiled analytic-gradient "
                    "restatement of the reference's TFP graph on all host cores"}
    print(json.dumps(line), flush=True)
    return 0


def reference_build_rate(cores, n_mats):
    """The reference's set-up route for one (dataset, component): its own `_build_matrices` restated (scipy kvp on n^2
    points, inner pinv, two GEMMs; magi_v2.py:774-823) + two SVD pseudo-inverses (tf.linalg.pinv, :126-128), n = 161,
    one matrix per thread at a time.  Returns (ms per matrix per core, matrices/s on all cores)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    from oracle import magi_oracle as mo
    I = np.linspace(0.0, 4.0, N_GRID)

    def one(k):
        t0 = time.perf_counter()
        C_d, m_d, K_d = mo.build_matrices(I, 0.01 + 0.001 * k, 0.15 + 0.01 * (k % 10))
        mo.tf_pinv(C_d); mo.tf_pinv(K_d)
        return time.perf_counter() - t0

    try:
        from threadpoolctl import threadpool_limits
        lim = threadpool_limits(limits=1)
    except ImportError:
        lim = None
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        per = list(ex.map(one, range(n_mats)))
    wall = time.perf_counter() - t0
    del lim
    return 1e3 * float(np.mean(per)), n_mats / wall


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(dev_index):
    """Best effort: run this rank (and allocate its pinned memory) on the CPUs local to its GPU (sysfs local_cpulist)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev_index)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus, {path})"
    except Exception as e:  # noqa: BLE001  (sysfs layout / permissions differ between boxes)
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def spot_check(model, band, n, consts_of, mats_of, X, s, tau, bt, outs, picks):
    """Post-hoc checker: the measured launch's outputs at `picks` [(dataset, chain)] against the C oracle on the
    matrices the device built.  consts_of(b) -> dict(mu, y, mask, N_ds, beta, LB); mats_of(b) -> (Cinv, m, Kinv) host."""
    import numpy as np
    from oracle import c_oracle as co
    from oracle import magi_oracle as mo
    lp, gX, gs, gt = outs
    worst, cache = 0.0, {}
    Dm = mo.MODELS[model].D
    for b, r in picks:
        if b not in cache:
            Cinv, m, Kinv = mats_of(b)
            k = consts_of(b)
            idx = np.where(k["mask"].reshape(-1) > 0)[0]
            oc = mo.PosteriorConstants(I=np.zeros((n, 1)), mu_ds=k["mu"], C_d_invs=Cinv, m_ds=m, K_d_invs=Kinv,
                                       N_ds=k["N_ds"], not_nan_idxs=idx, not_nan_cols=idx % Dm,
                                       y_tau_ds_observed=k["y"].reshape(-1)[idx], beta=float(k["beta"]),
                                       sigma_sqs_LB=k["LB"], f_vec=mo.MODELS[model].f_vec)
            cache = {b: co.COracle(oc, model, band=band)}
        lpo, go = cache[b].logpost_grad(mo.pack_state(X[b, r].cpu().numpy(), s[b, r].cpu().numpy(),
                                                      tau[b, r].cpu().numpy()), float(bt[b, r]))
        gg = mo.pack_state(gX[b, r].cpu().numpy(), gs[b, r].cpu().numpy(), gt[b, r].cpu().numpy())
        worst = max(worst, abs(float(lp[b, r]) - lpo) / abs(lpo), float(np.abs(gg - go).max() / np.abs(go).max()))
    return {"k": len(picks), "max_rel_err": worst, "tolerance": 1e-9, "ok": bool(worst <= 1e-9),
            "checker": "oracle/magi_oracle_c.c on device-built matrices"}


def other_configs(dev):
    """Evaluation throughput (magi_b200 log-posterior + gradient, inputs resident) on the other BASELINE.json
    configs, which are parity-test cases rather than the bench line: synthetic constants, device-built matrices,
    the evaluation path `PosteriorProblem.logpost_grad(path="auto")` picks, and a parity spot check against the oracle.
    Rank 0 only, a few seconds."""
    import numpy as np
    import torch
    from magi_v2_b200 import ops
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    res = []
    for name, model, D_, P_, n, B, R, band in (("config 2: 20 SEIR datasets", "seir4", 4, 3, 161, 20, 8, 80),
                                               ("config 3: SIRW n=321", "sirw", 4, 5, 321, 512, 8, None),
                                               ("config 5: Lorenz-96 n=1281", "lorenz96", 10, 1, 1281, 2, 64, None)):
        rng = np.random.default_rng(0)
        I = np.linspace(0, 4, n)
        phi1, phi2 = rng.uniform(0.01, 0.05, (B, D_)), rng.uniform(0.15, 0.3, (B, D_))
        C_, Cp, Cpp = ops.cov_build(T(I), T(phi1), T(phi2), 2.01, True)
        Cinv, m, Kinv, _, info = ops.factor_derive(C_, Cp, Cpp, -1 if band is None else band, 0.0)
        ok = int(info.abs().max()) == 0
        packed = ops.pack_matrices(Cinv, m, Kinv)
        spot_b = [0, B - 1]
        host_mats = {b: tuple(a[b].cpu().numpy() for a in (Cinv, m, Kinv)) for b in spot_b}
        del C_, Cp, Cpp, Cinv, m, Kinv
        mask = np.zeros((B, n, D_), dtype=np.uint8)
        mask[:, ::(n - 1) // 80] = 1
        y = rng.normal(0.3, 0.1, (B, n, D_)) * mask
        mu, N_ds = np.full((B, D_), 0.3), np.full((B, D_), 81.0)
        beta, LB = np.full(B, D_ * n / (81.0 * D_)), np.full((B, D_), 1e-6)
        prob = ops.PosteriorProblem(model, packed, mu=T(mu), y=T(y), mask=T(mask, torch.uint8), N_ds=T(N_ds),
                                    beta=T(beta), LB=T(LB), n=n, band=band)
        X, s = T(rng.normal(0.3, 0.05, (B, R, n, D_))), T(rng.normal(-6, 0.5, (B, R, D_)))
        tau, bt = T(rng.normal(0.5, 0.2, (B, R, P_))), T(np.full((B, R), 0.37))
        out = prob.logpost_grad_out(R)
        for _ in range(3):
            prob.logpost_grad(X, s, tau, bt, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            prob.logpost_grad(X, s, tau, bt, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        spot = spot_check(model, band, n, lambda b: dict(mu=mu[b], y=y[b], mask=mask[b], N_ds=N_ds[b], beta=beta[b],
                                                         LB=LB[b]), lambda b: host_mats[b], X, s, tau, bt, out,
                          [(0, 0), (B - 1, R - 1)])
        bytes_eval = 24.0 * D_ * n * n / R + 16.0 * (n * D_ + D_ + P_)
        res.append({"config": name, "model": model, "n_grid": n, "D": D_, "datasets": B, "chains_per_dataset": R,
                    "bandsize": band, "path": prob.eval_path(R), "ms_per_launch": ms,
                    "evals_per_s": B * R / (ms * 1e-3), "algorithmic_gb_per_s": bytes_eval * B * R / (ms * 1e-3) / 1e9,
                    "fp64_tflops": 8.0 * D_ * n * n * B * R / (ms * 1e-3) / 1e12, "factorisation_ok": ok,
                    "parity_spot": spot})
        del prob, packed, X, out, host_mats
    return res


def fp64_peaks(dev):
    """DFMA and DMMA throughput of this GPU, measured now (include/magi_b200_probe.h), TFLOP/s."""
    import ctypes as C

    import torch
    from magi_v2_b200 import _lib
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    threads, iters = 512, 20000
    blocks = sms * (2048 // threads)
    buf = torch.empty(blocks * threads, dtype=torch.float64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    res = {}
    for kind, name in ((0, "dfma"), (1, "dmma")):
        fl = C.c_double()
        best = 0.0
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(_lib.lib().magi_b200_probe_fp64(kind, iters, blocks, threads, C.c_void_p(buf.data_ptr()),
                                                       C.byref(fl), st), "probe_fp64")
            e1.record()
            torch.cuda.synchronize(dev)
            if rep:                                          # first launch = warm-up
                best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        res[name] = best
    return res


def build_section(dev, fp64_peak, cores, with_cpu):
    """cov_build + factor_derive for 2048 matrices of 161^2 (512 datasets x 4 components), CUDA events."""
    import numpy as np
    import torch
    from magi_v2_b200 import ops
    Bb, n = 512, N_GRID
    rng = np.random.default_rng(1)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    I, p1, p2 = T(np.linspace(0, 4, n)), T(rng.uniform(0.005, 0.05, (Bb, D))), T(rng.uniform(0.1, 0.4, (Bb, D)))

    def ev_time(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    ms_cov = ev_time(lambda: ops.cov_build(I, p1, p2, 2.01, True))
    C_, Cp, Cpp = ops.cov_build(I, p1, p2, 2.01, True)
    ms_fac = ev_time(lambda: ops.factor_derive(C_, Cp, Cpp, 80, 0.0))
    info = ops.factor_derive(C_, Cp, Cpp, 80, 0.0)[4]
    nmat = Bb * D
    flops = 6.0 * n ** 3 * nmat
    out = {"matrices": nmat, "n": n, "cov_build_ms": ms_cov, "cov_build_write_gb_per_s": 24.0 * n * n * nmat / (ms_cov * 1e-3) / 1e9,
           "factor_derive_ms": ms_fac, "factor_tflops": flops / (ms_fac * 1e-3) / 1e12,
           "factor_frac_of_fp64_peak": flops / (ms_fac * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
           "factorisation_ok": int(info.abs().max()) == 0,
           "matrices_per_s": nmat / ((ms_cov + ms_fac) * 1e-3)}
    if with_cpu:
        ms_ref, rate = reference_build_rate(cores, max(cores, 8))
        out["reference_route"] = {"ms_per_matrix_per_core": ms_ref, "matrices_per_s": rate, "cores": cores,
                                  "what": "restated _build_matrices (scipy kvp x3 on n^2 points, pinv, 2 GEMMs; "
                                          "magi_v2.py:774-823) + 2 SVD pseudo-inverses (:126-128), n = 161"}
        out["speedup_vs_reference_route"] = out["matrices_per_s"] / rate
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MAGI kernels have no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else "unchanged (single rank)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from magi_v2_b200 import synth
    from magi_v2_b200.parallel import gather_samples

    B, R, L = args.datasets, args.chains, args.leapfrog
    t_setup = time.perf_counter()
    prob, info, state, data = synth.sweep_problem(B, R, dev, seed0=rank * B, model="seir4", bandsize=80)
    assert int(info.abs().max()) == 0, "factorisation reported a non-positive-definite matrix"
    if os.environ.get("MAGI_BENCH_DENSE"):      # experiment: read the (banded) matrices as if dense
        prob.band = -1
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).pin_memory()
    hX, hs, ht = pin(state["X"]), pin(state["sig_pre"]), pin(state["th_pre"])
    hbt = pin(np.full((B, R), 0.37))
    X, s, tau, bt = (a.to(dev) for a in (hX, hs, ht, hbt))
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        """W warm-ups, then exactly `steps` calls bracketed by barrier+synchronize, CUDA events on the
        launching (current) stream; returns max-over-ranks milliseconds."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()

    # -- (1) device-resident evaluations: ONE launch of magi_b200_logpost_grad per step -----------
    out = prob.logpost_grad_out(R)
    ms_dev = timed(lambda: prob.logpost_grad(X, s, tau, bt, out=out), args.steps, max(3, args.warmup))
    evals_per_step = B * R * world
    value = evals_per_step * args.steps / (ms_dev * 1e-3)

    # -- (2) end to end: chain states in pinned host blocks, one H2D + launch + D2H per dataset chunk -----------------
    hp = prob.host_pipeline(R, n_chunks=args.e2e_chunks, n_streams=args.e2e_streams)
    hp.fill(hX, hs, ht, hbt)
    ms_e2e = timed(lambda: hp.run(wait=True), args.steps, max(3, args.warmup))
    e2e_value = evals_per_step * args.steps / (ms_e2e * 1e-3)
    h2d, d2h = hp.h2d_bytes, hp.d2h_bytes
    # (2b) the same with two buffer sets: batch k+1 is uploaded and evaluated while batch k is still being downloaded
    # (independent batches; a host-side sampler, whose next inputs depend on this step's outputs, sees (2))
    from magi_v2_b200.ops import HostPipeline
    hp2 = HostPipeline(prob, R, args.e2e_chunks, args.e2e_streams, ws_slot0=1 + args.e2e_streams)
    hp2.fill(hX, hs, ht, hbt)
    pipes, side = [hp, hp2], [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def e2e_double_buffered():
        for k in range(args.steps):
            p = pipes[k & 1]
            if k >= 2:
                p.done.synchronize()          # this buffer set's previous results are complete (a user reads them here)
            with torch.cuda.stream(side[k & 1]):
                p.run(wait=False)
        for p in pipes:
            p.done.synchronize()

    e2e_double_buffered()
    barrier()
    t0 = time.perf_counter()
    e2e_double_buffered()
    ms_db = (time.perf_counter() - t0) * 1e3
    ms_db = max_over_ranks(ms_db)
    if clocks:
        clk = clocks.stop()
    # the host path returns what the device path returns
    e2e_same = bool(torch.equal(hp.gather()[0], out[0].cpu()))

    # -- (3) fused HMC sampler ---------------------------------------------------------------------
    hmc, allgather = None, None
    if not args.no_hmc:
        eps = torch.full((B, R), 2e-4, dtype=torch.float64, device=dev)
        da = torch.zeros((B, R, 4), dtype=torch.float64, device=dev)
        Xh, sh, th = X.clone(), s.clone(), tau.clone()
        it = [0]

        def hmc_step():
            prob.hmc_run_(Xh, sh, th, eps, da, n_iter=args.hmc_iters, n_leapfrog=L, iter0=it[0], num_adapt=0,
                          seed=1 + rank, chain_id0=rank * B * R, fixed_beta_temp=0.37)
            it[0] += args.hmc_iters

        hsteps = max(2, args.steps // 4)
        ms_hmc = timed(hmc_step, hsteps, 3)
        samples = B * R * world * args.hmc_iters * hsteps
        # end to end: chain states up from pinned host memory, transitions, theta/sigma samples and the
        # final states back down; with world > 1 the theta samples are all-gathered (the one collective)
        last, pinned = {}, {}

        def to_pinned(name, t):
            """device -> a pinned host buffer kept across calls (a pageable `.cpu()` runs at a third of the link)"""
            if name not in pinned or pinned[name].shape != t.shape:
                pinned[name] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            pinned[name].copy_(t, non_blocking=True)
            return pinned[name]

        def hmc_e2e():
            Xe, se, te = hX.to(dev, non_blocking=True), hs.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
            o = prob.hmc_run_(Xe, se, te, eps, da, n_iter=args.hmc_iters, n_leapfrog=L, iter0=0, num_adapt=0,
                              seed=1 + rank, chain_id0=rank * B * R, fixed_beta_temp=0.37)
            last["th"] = o["thetas_samps"]
            ths = gather_samples(o["thetas_samps"]) if world > 1 else o["thetas_samps"]
            res = to_pinned("th", ths), to_pinned("sig", o["sigma_sqs_samps"]), to_pinned("X", Xe)
            torch.cuda.synchronize()          # the host buffers are complete
            return res

        ms_hmc_e2e = timed(hmc_e2e, hsteps, 3)
        hmc = {"samples_per_s": samples / (ms_hmc * 1e-3), "n_leapfrog": L, "transitions_per_launch": args.hmc_iters,
               "evals_per_s_inside_sampler": samples * (L + 0.0) / (ms_hmc * 1e-3),
               "e2e_samples_per_s": samples / (ms_hmc_e2e * 1e-3), "ms_per_launch": ms_hmc / hsteps,
               "accept_rate_note": "step size 2e-4, fixed beta_temp 0.37, no adaptation"}
        if world > 1:
            ms_ag = timed(lambda: gather_samples(last["th"]), 5, 2)
            nbytes = last["th"].numel() * 8 * world
            allgather = {"what": "all_gather_into_tensor of the theta samples (the only collective, SURVEY.md 8e)",
                         "bytes_gathered_per_rank": nbytes, "ms": ms_ag / 5, "gb_per_s": nbytes / (ms_ag / 5 * 1e-3) / 1e9}

    # -- (4) NUTS (the reference's sampler): tree building in nuts.py, one logpost_grad launch per leapfrog --------
    nuts_res = None
    if not args.no_nuts:
        from magi_v2_b200 import nuts
        zN = nuts.pack_state(X, s, tau)
        epsN = torch.full((B * R,), 2e-4, dtype=torch.float64, device=dev)
        daN = torch.zeros((B * R, 4), dtype=torch.float64, device=dev)
        ids = torch.arange(rank * B * R, (rank + 1) * B * R, dtype=torch.int64, device=dev)
        eng = nuts.FusedLeafEngine(prob, R)
        itn, leaves = [0], [0]

        def nuts_step():
            o = nuts.nuts_run_(zN, epsN, daN, None, n_iter=1, iter0=itn[0], num_adapt=0, fixed_beta_temp=0.37,
                               seed=1 + rank, chain_ids=ids, max_tree_depth=args.nuts_depth, leaf_engine=eng)
            itn[0] += 1
            leaves[0] += int(o["n_leapfrog"].sum())

        nuts_step()
        leaves[0] = 0
        nsteps = 2
        ms_nuts = timed(nuts_step, nsteps, 1)
        launches_timed = nsteps * (1 << args.nuts_depth)        # 2^depth - 1 leaves + the starting point
        tot = torch.tensor([leaves[0] * nsteps / (nsteps + 1.0)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        nuts_res = {"transitions_per_s": B * R * world * nsteps / (ms_nuts * 1e-3), "max_tree_depth": args.nuts_depth,
                    "leapfrogs_per_s": float(tot.item()) / (ms_nuts * 1e-3),
                    "mean_leapfrogs_per_transition": float(tot.item()) / (B * R * world * nsteps),
                    "ms_per_lockstep_leapfrog": ms_nuts / max(launches_timed, 1),
                    "note": "step size 2e-4, fixed beta_temp 0.37; every chain of every dataset advances in lock-step, "
                            "per leaf: magi_b200_nuts_leaf_pre, magi_b200_logpost_grad, magi_b200_nuts_leaf_post"}
        del zN

    e2e_rank_ms = None
    if world > 1:                                   # per-rank e2e time (the max is what `e2e.value` uses)
        t = torch.tensor([ms_e2e / args.steps], dtype=torch.float64, device=dev)
        allms = [torch.zeros_like(t) for _ in range(world)]
        # every rank measured the max already; re-measure locally without the cross-rank max
        for _ in range(2):
            hp.run(wait=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            hp.run(wait=True)
        t[0] = (time.perf_counter() - t0) / 5 * 1e3
        dist.all_gather(allms, t)
        e2e_rank_ms = [float(a.item()) for a in allms]

    others = None
    if rank == 0 and not args.no_other_configs:
        others = other_configs(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # -- parity of the measured launch (rank 0): k random (dataset, chain) outputs against the oracle ---------------------
    from magi_v2_b200 import ops
    rng = np.random.default_rng(2026)
    picks = [(int(b), int(r)) for b, r in zip(rng.integers(0, B, 8), rng.integers(0, R, 8))]
    c = data["consts"]
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)

    def mats_of(b):
        C_, Cp, Cpp = ops.cov_build(T(c["I"]), T(data["phi1"][b:b + 1]), T(data["phi2"][b:b + 1]), 2.01, True)
        Cinv, m, Kinv, _, _ = ops.factor_derive(C_, Cp, Cpp, 80, 0.0)
        return Cinv[0].cpu().numpy(), m[0].cpu().numpy(), Kinv[0].cpu().numpy()

    prob.logpost_grad(X, s, tau, bt, out=out)
    torch.cuda.synchronize()
    parity = spot_check("seir4", 80, N_GRID, lambda b: dict(mu=c["mu"][b], y=c["y"][b], mask=c["mask"][b], N_ds=c["N_ds"][b],
                                                            beta=c["beta"][b], LB=data["LB"][b]), mats_of, X, s, tau, bt,
                        out, picks)
    parity["picks"] = picks
    parity["host_path_equals_device_path"] = e2e_same

    # -- roofline of the dominant kernel (logpost_grad_fast_kernel<Seir4, 168>) -----------------------------
    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except (OSError, KeyError, ValueError):
        peak = 6650.0
        peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    fp = fp64_peaks(dev)
    bytes_per_eval = 24.0 * D * N_GRID * N_GRID / R + 16.0 * S_STATE
    launch_ms = ms_dev / args.steps
    achieved = bytes_per_eval * B * R / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("logpost_grad_fast_kernel_bytes_per_launch_4096_datasets")
            traffic_src = tj.get("source")
            if traffic is not None and B != 4096:
                traffic = traffic * B / 4096.0
        except (OSError, ValueError):
            traffic = None
    tiles_read = 331.0 / 441.0        # band 80 of n = 161: non-zero 8x8 tiles of the 21 x 21 tile grid
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "logpost_grad_fast_kernel<Seir4, 168>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_eval * B * R,
                "designed_bytes_per_launch": (24.0 * D * 168 * 168 * tiles_read / R + 16.0 * S_STATE) * B * R,
                "designed_bytes_note": "what the kernel is built to read: only the tiles inside the band "
                                       "(331 of 441 per matrix) of the 168-padded matrices, once per 8 chains",
                "fp64_flops_per_launch": 8.0 * D * N_GRID * N_GRID * B * R,
                "fp64_tflops_achieved": 8.0 * D * N_GRID * N_GRID * B * R / (launch_ms * 1e-3) / 1e12,
                "fp64_peak_tflops_measured": {"dmma": fp["dmma"], "dfma": fp["dfma"],
                                              "how": "magi_b200_probe_fp64, this run, CUDA events"}}

    cores = host_cores()
    build = None
    if not args.no_build:
        build = build_section(dev, fp["dmma"], cores, with_cpu=(world == 1 and not args.no_cpu_baseline))

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6",
                            "--warmup", "1", "--datasets", str(B), "--chains", str(R)], capture_output=True,
                           text=True, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                cpu_baseline = json.loads(ln)["cpu_baseline"]
        if cpu_baseline is None:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                            "sample": "failed: " + r.stderr[-300:]}

    step_s = ms_e2e / args.steps * 1e-3
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "chunks": hp.n_chunks, "streams": len(hp.streams),
                    "copies_per_step": 2 * hp.n_chunks, "pcie_gb_per_s_each_way": h2d / step_s / 1e9,
                    "per_rank_ms_per_step": e2e_rank_ms, "cpu_affinity": affinity,
                    "double_buffered": {"value": evals_per_step * args.steps / (ms_db * 1e-3), "unit": UNIT,
                                        "ms_per_step": ms_db / args.steps,
                                        "note": "two buffer sets alternated (HostPipeline x 2): the upload and evaluation "
                                                "of batch k+1 overlap the download of batch k; host wall clock, max over "
                                                "ranks; every step still copies its inputs in and its results out"}},
            "gpu_launches": args.steps, "clocks": clk, "roofline": roofline, "parity_spot": parity,
            "cpu_baseline": cpu_baseline, "build": build, "hmc": hmc, "nuts": nuts_res, "allgather": allgather,
            "other_configs": others, "oracle_pin": ORACLE_PIN, "setup_s": t_setup}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
