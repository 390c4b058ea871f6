"""The reference's vignette (vignette.ipynb: SEIR seed 0, E/I/R observed, discretization 1, band 80, 1000 + 1000 NUTS
iterations) replayed END TO END on the CPU oracle: hyper-parameter fit (oracle.init_oracle, magi_v2.py:538-691), theta
initialisation as written (:132-179), banding, spline smoothing, then the sampler stack of :357-371 on the C oracle
(oracle/magi_oracle_c.c), several independent chains in parallel processes.

TEST INFRASTRUCTURE (oracle/__init__.py).  Writes a JSON summary; profiles/r02_vignette.md quotes it.

    python -m oracle.vignette_study --chains 8 --results 1000 --burnin 1000 --layout reference --out /tmp/v.json
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import time

import numpy as np

from . import c_oracle as co
from . import init_oracle as io
from . import magi_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def vignette_data(which: int = 0):
    g = np.load(os.path.join(GOLDEN, "seir_datasets.npz"))
    X = g["X_obs"][which][:, 1:].copy()
    X[X < 0.0] = 0.0                                   # vignette.ipynb:112-113
    return g["ts_obs"], X


def split_rhat(x):
    """x [chains, draws] -> split-R-hat (Gelman et al., BDA3)."""
    C, N = x.shape
    h = N // 2
    s = np.concatenate([x[:, :h], x[:, h:2 * h]], axis=0)
    W = s.var(axis=1, ddof=1).mean()
    B = h * s.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((h - 1) / h * W + B / h) / W))


def _run_chain(args):
    fit, chain, opt = args
    c = fit["constants"]
    O = co.COracle(c, "seir3", band=opt["band"])
    z0 = mo.pack_state(*mo.initial_state(fit["Xhat_init"], fit["sigma_sqs_init"], fit["thetas_init"], c.sigma_sqs_LB))
    n_iter = opt["burnin"] + opt["results"]
    r = O.nuts_chain(z0, n_iter, eps0=0.1, seed=opt["seed"], chain_id=chain,
                     num_adaptation_steps=int(0.8 * opt["burnin"]), fixed_beta_temp=opt["fixed_bt"],
                     max_tree_depth=opt["depth"], cached_lp=opt["cached"], store_z=False)
    r.pop("z")
    return r


def summarize(runs, fit, opt):
    D, LB = 3, fit["constants"].sigma_sqs_LB
    nb = opt["burnin"]
    sp = lambda v: np.logaddexp(0.0, v)
    tails = np.stack([r["tail"] for r in runs])                     # [chains, iters, D + P]
    th = sp(tails[:, nb:, D:])                                      # thetas_samps (:419)
    sg = sp(tails[:, nb:, :D]) + LB                                 # sigma_sqs_samps (:418)
    per_chain = th.mean(axis=1)
    out = dict(
        options=opt,
        phi1s=fit["phi1s"].tolist(), phi2s=fit["phi2s"].tolist(), sigma_sqs_init=fit["sigma_sqs_init"].tolist(),
        thetas_init=fit["thetas_init"].tolist(),
        theta_mean=th.mean(axis=(0, 1)).tolist(), theta_sd=th.reshape(-1, 3).std(axis=0).tolist(),
        theta_mean_per_chain=per_chain.tolist(),
        theta_mc_se_between_chains=(per_chain.std(axis=0, ddof=1) / np.sqrt(len(runs))).tolist() if len(runs) > 1 else None,
        theta_split_rhat=[split_rhat(th[:, :, k]) for k in range(3)] if len(runs) > 1 else None,
        sigma_sq_mean=sg.mean(axis=(0, 1)).tolist(),
        theta_at_burnin_end=sp(tails[:, nb - 1, D:]).tolist() if nb > 0 else None,
        theta_trace_every_100=sp(tails[:, ::100, D:]).tolist(),
        step_size_final=[float(r["step_size"][-1]) for r in runs],
        mean_leapfrogs=float(np.mean([r["leapfrogs"].mean() for r in runs])),
        mean_depth=float(np.mean([r["depth"].mean() for r in runs])),
        mean_accept=float(np.mean([r["accept"][nb:].mean() for r in runs])),
        lp_last=[float(r["lp"][-1]) for r in runs],
        vignette_printed=[5.831, 0.565, 1.77], truth=[6.0, 0.6, 1.8],
    )
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--results", type=int, default=1000)
    ap.add_argument("--burnin", type=int, default=1000)
    ap.add_argument("--layout", default="reference", choices=["reference", "transpose"])
    ap.add_argument("--cached", type=int, default=1)
    ap.add_argument("--fixed-bt", type=float, default=None)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--hparams", default=None, help="npz with phi1s/phi2s/sigma_sqs (skips the 1000-step fit)")
    ap.add_argument("--thetas-init", default=None, help="comma-separated override of thetas_init")
    ap.add_argument("--out", default=None)
    ap.add_argument("--golden", action="store_true",
                    help="write tests/golden/vignette_chains.npz: the reference's settings, intended theta start, "
                         "hyper-parameters of tests/golden/vignette_fit.npz")
    a = ap.parse_args()
    ts, X = vignette_data()
    if a.golden:
        a.layout, a.cached, a.fixed_bt, a.burnin, a.results, a.depth = "transpose", 1, None, 1000, 1000, 10
        a.hparams = os.path.join(GOLDEN, "vignette_fit.npz")
    hp = dict(np.load(a.hparams)) if a.hparams else None
    if hp is not None:
        hp = {k: hp[k] for k in ("phi1s", "phi2s", "sigma_sqs")}
    t0 = time.time()
    fit = io.initial_fit(ts, X, 1, 80, mo.f_seir3, 3, hparams=hp, theta_layout=a.layout)
    if a.thetas_init:
        fit["thetas_init"] = np.array([float(v) for v in a.thetas_init.split(",")])
    print(f"initial_fit {time.time() - t0:.1f} s: phi1 {fit['phi1s']} phi2 {fit['phi2s']} sigma_sq {fit['sigma_sqs_init']} "
          f"thetas_init {fit['thetas_init']}", flush=True)
    opt = dict(burnin=a.burnin, results=a.results, fixed_bt=a.fixed_bt, depth=a.depth, cached=a.cached, seed=a.seed,
               band=80, layout=a.layout, chains=a.chains)
    t0 = time.time()
    with mp.Pool(min(a.chains, os.cpu_count())) as pool:
        runs = pool.map(_run_chain, [(fit, ch, opt) for ch in range(a.chains)])
    out = summarize(runs, fit, opt)
    out["sampling_seconds"] = time.time() - t0
    print(json.dumps({k: out[k] for k in ("theta_mean", "theta_sd", "theta_mc_se_between_chains", "theta_split_rhat",
                                          "sigma_sq_mean", "step_size_final", "mean_leapfrogs", "mean_accept",
                                          "theta_at_burnin_end", "sampling_seconds")}, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    if a.golden:
        D, nb = 3, a.burnin
        tails = np.stack([r["tail"] for r in runs])[:, nb:]
        sp = lambda v: np.logaddexp(0.0, v)
        np.savez_compressed(os.path.join(GOLDEN, "vignette_chains.npz"),
                            theta_chain_means=sp(tails[:, :, D:]).mean(axis=1),
                            log_sigma_sq_chain_means=np.log(sp(tails[:, :, :D]) + fit["constants"].sigma_sqs_LB).mean(axis=1),
                            step_size_final=np.array([r["step_size"][-1] for r in runs]),
                            mean_leapfrogs=np.array([r["leapfrogs"].mean() for r in runs]),
                            settings=np.array([a.burnin, a.results, a.depth, a.cached, a.seed, a.chains]))


if __name__ == "__main__":
    main()
