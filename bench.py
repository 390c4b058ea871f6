#!/usr/bin/env python
"""Benchmark of the MAGI posterior-evaluation hot path on B200 (driver contract: see README/DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (restated)

Workload (BASELINE.json configs[3]): synthetic SEIR-shaped sweep,
B = 4096 datasets x R = 8 chains PER GPU (weak scaling: every rank owns its own 4096 datasets, no
data-path collective), n = 161 grid points, D = 4 components, P = 3 parameters, band 80.
A "step" is one log-posterior + analytic-gradient evaluation of all B*R chains (one launch of
`magi_b200_logpost_grad`); `value` = evaluations/s with inputs resident in HBM; `e2e` = the same
through the host-buffer entry point (pinned host -> device copies of the chain states and device ->
host copies of lp and all gradients inside the timed region).  The `hmc` object reports HMC
transitions/s of the fused sampler kernel (L leapfrog steps per transition) the same two ways; `nuts` the
transitions/s and leapfrogs/s of the reference's No-U-Turn sampler on the same chains (magi_v2_b200/nuts.py on the
fused leaf kernels); `other_configs` (rank 0, a few seconds) the evaluation throughput of BASELINE configs 2, 3, 5.

The reference arm times the op-for-op restatement of the reference's TFP graph (oracle/, torch CPU
FP64 + autograd, one chain per call as the reference does) on all host cores; TensorFlow-Probability
itself is not installable in this image.  The oracle is executed by this file only there and in the
`cpu_baseline` leg of the CUDA arm (a bounded sample on rank 0); the measured CUDA path never imports it."""
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
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short evaluation-throughput lines of BASELINE configs 2, 3 and 5")
    ap.add_argument("--nuts-depth", type=int, default=4, help="max_tree_depth of the timed NUTS transitions")
    ap.add_argument("--cpu-evals-per-worker", type=int, default=24)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "synthetic SEIR4 sweep (BASELINE configs[3]): datasets x chains of independent MAGI posteriors",
            "datasets_per_gpu": args.datasets, "chains_per_dataset": args.chains, "n_grid": N_GRID, "D": D, "P": P,
            "bandsize": 80, "sharding": f"datasets over {n_gpus} rank(s), no data-path collective",
            "l2": "inputs (packed matrices, %.1f GB per GPU) exceed the 126 MB L2; no flush needed"
                  % (args.datasets * D * 3 * 168 * 168 * 8 / 1e9)}


# ------------------------------------------------------------------------------------------------
# reference arm: the restated reference CPU path on all host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(seed0):
    import numpy as np
    import torch
    torch.set_num_threads(1)
    try:
        from threadpoolctl import threadpool_limits
        _W["blas_limit"] = threadpool_limits(limits=1)       # one BLAS thread per worker process
    except ImportError:
        pass
    from magi_v2_b200 import synth
    from oracle import magi_oracle as mo
    data = synth.seir_sweep(1, seed0=seed0, model="seir4")
    rng = np.random.default_rng(seed0)
    phi1, phi2 = rng.uniform(0.005, 0.05, D), rng.uniform(0.1, 0.4, D)
    c = mo.make_constants(data["ts_obs"], data["X_obs"][0], 1, phi1, phi2, 80, mo.f_seir4)
    y, mask = c.dense_y_mask()
    Xhat = mo.linear_interpolate(np.where(mask > 0, y, np.nan))
    _W.update(mo=mo, c=c, Xhat=Xhat, rng=rng)
    return True


def _worker_eval(n_evals):
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    import numpy  # noqa: F401  (imported before forking so that the workers share the loaded modules)
    import scipy.special  # noqa: F401
    import torch
    torch.set_num_threads(1)
    from magi_v2_b200 import synth  # noqa: F401
    from oracle import magi_oracle  # noqa: F401
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    workers = max(1, cores)
    per = args.cpu_evals_per_worker
    ctx = mp.get_context("fork")
    # one process per core, each holding its own dataset's constants (built by the oracle's
    # restatement of the reference route: scipy kvp + SVD pseudo-inverses, magi_v2.py:774-823, :126-128)
    pools = [ctx.Pool(1, initializer=_worker_init, initargs=(1000 + w,)) for w in range(workers)]

    def step():
        rs = [p.apply_async(_worker_eval, (per,)) for p in pools]
        return sum(r.get() for r in rs)

    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for p in pools:
        p.close()
    evals = args.steps * workers * per
    value = evals / dt
    sample = (f"{workers} worker processes x {per} evaluations per step x {args.steps} steps; each evaluation = "
              "restated unnormalized_log_prob (magi_v2.py:308-348) + torch autograd, one chain per call, "
              "SEIR4 n=161 D=4 band 80, torch CPU FP64, 1 thread per worker")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "TensorFlow-Probability is not installable in this image; this is the oracle's op-for-op "
                    "restatement of the reference's TFP graph (kind=port), on all host cores"}
    print(json.dumps(line), flush=True)
    return 0


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


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def other_configs(dev):
    """Evaluation throughput (magi_b200 log-posterior + gradient, inputs resident) on the other BASELINE.json
    configs, which are parity-test cases rather than the bench line: synthetic constants, device-built matrices,
    the evaluation path `PosteriorProblem.logpost_grad(path="auto")` picks.  Rank 0 only, a few seconds."""
    import numpy as np
    import torch
    from magi_v2_b200 import ops
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    res = []
    for name, model, D, P, n, B, R, band in (("config 2: 20 SEIR datasets", "seir4", 4, 3, 161, 20, 8, 80),
                                             ("config 3: SIRW n=321", "sirw", 4, 5, 321, 512, 8, None),
                                             ("config 5: Lorenz-96 n=1281", "lorenz96", 10, 1, 1281, 2, 64, None)):
        rng = np.random.default_rng(0)
        I = np.linspace(0, 4, n)
        phi1, phi2 = rng.uniform(0.01, 0.05, (B, D)), rng.uniform(0.15, 0.3, (B, D))
        C_, Cp, Cpp = ops.cov_build(T(I), T(phi1), T(phi2), 2.01, True)
        Cinv, m, Kinv, _, info = ops.factor_derive(C_, Cp, Cpp, -1 if band is None else band, 0.0)
        ok = int(info.abs().max()) == 0
        packed = ops.pack_matrices(Cinv, m, Kinv)
        del C_, Cp, Cpp, Cinv, m, Kinv
        mask = np.zeros((B, n, D), dtype=np.uint8)
        mask[:, ::(n - 1) // 80] = 1
        y = rng.normal(0.3, 0.1, (B, n, D)) * mask
        prob = ops.PosteriorProblem(model, packed, mu=T(np.full((B, D), 0.3)), y=T(y), mask=T(mask, torch.uint8),
                                    N_ds=T(np.full((B, D), 81.0)), beta=T(np.full(B, D * n / (81.0 * D))),
                                    LB=T(np.full((B, D), 1e-6)), n=n, band=band)
        X, s = T(rng.normal(0.3, 0.05, (B, R, n, D))), T(rng.normal(-6, 0.5, (B, R, D)))
        tau, bt = T(rng.normal(0.5, 0.2, (B, R, P))), T(np.full((B, R), 0.37))
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
        bytes_eval = 24.0 * D * n * n / R + 16.0 * (n * D + D + P)
        res.append({"config": name, "model": model, "n_grid": n, "D": D, "datasets": B, "chains_per_dataset": R,
                    "bandsize": band, "path": prob.eval_path(R), "ms_per_launch": ms,
                    "evals_per_s": B * R / (ms * 1e-3), "algorithmic_gb_per_s": bytes_eval * B * R / (ms * 1e-3) / 1e9,
                    "fp64_tflops": 8.0 * D * n * n * B * R / (ms * 1e-3) / 1e12, "factorisation_ok": ok,
                    "finite": bool(torch.isfinite(out[0]).all())})
        del prob, packed, X, out
    return res


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

    # -- (2) end to end through the host-buffer entry point ---------------------------------------
    hout = prob.logpost_grad_host_out(R)
    e2e_chunks = int(os.environ.get("MAGI_E2E_CHUNKS", "16"))
    e2e_streams = int(os.environ.get("MAGI_E2E_STREAMS", "4"))
    ms_e2e = timed(lambda: prob.logpost_grad_host(hX, hs, ht, hbt, out=hout, n_chunks=e2e_chunks,
                                                  n_streams=e2e_streams), args.steps, max(3, args.warmup))
    e2e_value = evals_per_step * args.steps / (ms_e2e * 1e-3)
    h2d = (hX.numel() + hs.numel() + ht.numel() + hbt.numel()) * 8
    d2h = sum(t.numel() for t in hout) * 8
    if clocks:
        clk = clocks.stop()

    # -- (3) fused HMC sampler ---------------------------------------------------------------------
    hmc = None
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
        def hmc_e2e():
            Xe, se, te = hX.to(dev, non_blocking=True), hs.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
            o = prob.hmc_run_(Xe, se, te, eps, da, n_iter=args.hmc_iters, n_leapfrog=L, iter0=0, num_adapt=0,
                              seed=1 + rank, chain_id0=rank * B * R, fixed_beta_temp=0.37)
            ths = gather_samples(o["thetas_samps"]) if world > 1 else o["thetas_samps"]
            return ths.to("cpu", non_blocking=False), o["sigma_sqs_samps"].cpu(), Xe.cpu()

        ms_hmc_e2e = timed(hmc_e2e, hsteps, 3)
        hmc = {"samples_per_s": samples / (ms_hmc * 1e-3), "n_leapfrog": L, "transitions_per_launch": args.hmc_iters,
               "evals_per_s_inside_sampler": samples * (L + 0.0) / (ms_hmc * 1e-3),
               "e2e_samples_per_s": samples / (ms_hmc_e2e * 1e-3), "ms_per_launch": ms_hmc / hsteps,
               "accept_rate_note": "step size 2e-4, fixed beta_temp 0.37, no adaptation"}

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

    others = None
    if rank == 0 and not args.no_other_configs:
        others = other_configs(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # -- roofline of the dominant kernel (logpost_grad_kernel<Seir4>) -----------------------------
    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except (OSError, KeyError, ValueError):
        peak = 6650.0
        peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    bytes_per_eval = 24.0 * D * N_GRID * N_GRID / R + 16.0 * S_STATE
    launch_ms = ms_dev / args.steps
    achieved = bytes_per_eval * B * R / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("logpost_grad_fast_kernel_bytes_per_launch_4096_datasets")
        except (OSError, ValueError):
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "logpost_grad_fast_kernel<Seir4, 168>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_eval * B * R,
                "fp64_flops_per_launch": 8.0 * D * N_GRID * N_GRID * B * R,
                "fp64_tflops_achieved": 8.0 * D * N_GRID * N_GRID * B * R / (launch_ms * 1e-3) / 1e12,
                "fp64_peak_tflops_measured": 37.2}

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

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": args.steps, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "hmc": hmc, "nuts": nuts_res, "other_configs": others, "setup_s": t_setup}
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
