"""Fixed-length HMC with one evaluation launch per leapfrog step -- the counterpart of the fused `magi_b200_hmc_run`
kernel for the shapes where the wide evaluation (`magi_b200_logpost_grad_wide`, include/magi_b200_wide.h) is the faster
one: few datasets, or grids with np > 168.  Same algorithm and the same draws as the fused kernel and as
`oracle.hmc_chain` (TFP SimpleLeapfrogIntegrator, identity mass; momenta and the accept uniform from the Philox
stream of csrc/rng.cuh keyed by (global chain id, global iteration); per-chain dual averaging; the reference's
temperature schedule, magi_v2.py:833-835, :855), so the two samplers are interchangeable and are checked against the
oracle draw for draw.  The arithmetic is in the CUDA operators; this file is the control flow."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import nuts as _nuts

RNG_ACCEPT = 1


def hmc_run_host_(prob, X, sig_pre, th_pre, eps, da_state, *, n_iter: int, n_leapfrog: int, iter0: int = 0,
                  num_adapt: int = 0, accum_from: int = 0, min_temp: float = 0.1, fixed_beta_temp: float = 0.0,
                  target_accept: float = 0.75, seed: int = 0, chain_id0: int = 0, keep_theta=True, keep_sigma=True,
                  keep_X=False, X_sum: Optional[torch.Tensor] = None, X_sumsq: Optional[torch.Tensor] = None):
    """Same contract as `PosteriorProblem.hmc_run_` (in place on X, sig_pre, th_pre, eps, da_state)."""
    from ._lib import check, lib
    B, R, n, D, P = prob.B, X.shape[1], prob.n, prob.D, prob.P
    Cn, nD = B * R, n * D
    dev, f64 = prob.device, torch.float64
    mk = lambda *s: torch.empty(s, dtype=f64, device=dev)
    th_s = mk(n_iter, B, R, P) if keep_theta else None
    sg_s = mk(n_iter, B, R, D) if keep_sigma else None
    X_s = mk(n_iter, B, R, n, D) if keep_X else None
    acc, lpt = mk(n_iter, B, R), mk(n_iter, B, R)
    eng = _nuts.FusedLeafEngine(prob, R)                      # momentum / uniform kernels and the evaluation buffers
    ids = torch.arange(chain_id0, chain_id0 + Cn, dtype=torch.int64, device=dev)
    z = _nuts.pack_state(X, sig_pre, th_pre)
    e, da = eps.view(Cn), da_state.view(Cn, 4)
    LB = prob.LB[:, None, :]
    ke0, ke1 = mk(Cn), mk(Cn)
    S = z.shape[1]
    st_ptr = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None

    def kick(zc, pc, gc, ck, drift, energy):
        check(lib().magi_b200_hmc_kick_drift(Cn, S, ptr(zc), ptr(pc), ptr(gc), ptr(e), float(ck), int(drift), ptr(energy),
                                             st_ptr), "hmc_kick_drift")

    for it in range(n_iter):
        g_it = iter0 + it
        bt = float(fixed_beta_temp) if fixed_beta_temp > 0.0 else max(1.0 / math.log(g_it + 2.0), min_temp)
        eng.set_beta_temp(bt)
        lp0, g0 = eng.value_and_grad(z)
        p0 = eng.momentum(seed, ids, g_it)
        u, _ = eng.uniforms(seed, ids, g_it, RNG_ACCEPT, 0, 1)
        zc, pc, gc, lpc = z.clone(), p0.clone(), g0, lp0
        # TFP SimpleLeapfrogIntegrator on the device: one fused kick (+ drift) launch between evaluations -- the two half
        # kicks of consecutive steps are one full kick -- and the kinetic energies from the same kernel
        kick(zc, pc, gc, 0.0, False, ke0)                      # energy of the drawn momentum
        for st in range(n_leapfrog):
            kick(zc, pc, gc, 0.5 if st == 0 else 1.0, True, None)
            lpc, gc = eng.value_and_grad(zc)
        if n_leapfrog > 0:
            kick(zc, pc, gc, 0.5, False, ke1)
        else:
            ke1.copy_(ke0)
        dH = (-lpc + ke1) - (-lp0 + ke0)
        ap = torch.where(torch.isfinite(dH), torch.exp(torch.clamp(-dH, max=0.0)), torch.zeros_like(dH))
        accept = u[:, 0] < ap
        z.copy_(torch.where(accept[:, None], zc, z))
        _nuts.dual_averaging_update_(e, da, ap, num_adapt, target_accept)
        acc[it] = ap.view(B, R)
        lpt[it] = torch.where(accept, lpc, lp0).view(B, R)
        zz = z.view(B, R, -1)
        if keep_theta:
            th_s[it] = torch.nn.functional.softplus(zz[..., nD + D:], threshold=700.0)
        if keep_sigma:
            sg_s[it] = torch.nn.functional.softplus(zz[..., nD:nD + D], threshold=700.0) + LB
        if keep_X:
            X_s[it] = zz[..., :nD].reshape(B, R, n, D)
        if g_it >= accum_from and (X_sum is not None or X_sumsq is not None):
            xv = zz[..., :nD].reshape(B, R, n, D)
            if X_sum is not None:
                X_sum.add_(xv)
            if X_sumsq is not None:
                X_sumsq.addcmul_(xv, xv)
    zz = z.view(B, R, -1)
    X.copy_(zz[..., :nD].reshape(B, R, n, D)); sig_pre.copy_(zz[..., nD:nD + D]); th_pre.copy_(zz[..., nD + D:])
    return {"thetas_samps": th_s, "sigma_sqs_samps": sg_s, "X_samps": X_s, "accept_prob": acc, "lp": lpt}
