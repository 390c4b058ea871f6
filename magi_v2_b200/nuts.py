"""Batched No-U-Turn sampler on top of the device log-posterior + gradient operator.

The reference samples with ``tfp.mcmc.NoUTurnSampler(step_size=0.1)`` inside
``DualAveragingStepSizeAdaptation(target_accept_prob=0.75, num_adaptation_steps=0.8 * burn-in)`` inside its own
``LogAnnealedNUTS`` wrapper (magi_v2.py:357-371, :838-889), and hands NUTS one callable:
``target_log_prob_fn(X, sigma_sqs_pre, thetas_pre)`` (:361-364, :866-869) whose value and gradient TFP asks for once per
leapfrog step.  That callable is the drop-in boundary: here it is ``PosteriorProblem.logpost_grad`` (one launch of the
fused CUDA kernel for ALL chains of ALL datasets), and this module is the tree-building around it -- the batched,
iterative formulation TFP itself uses (all chains advance in lock-step, finished chains are masked, the U-turn checks
of the dyadic sub-trees use a checkpoint memory of max_tree_depth momenta and momentum sums instead of recursion).

Algorithm (restated from the published one; tensorflow-probability==0.24.0 is not installable here): multinomial
NUTS with the generalised U-turn criterion -- leaves weighted by exp(H0 - H); uniform progressive sampling inside a
new subtree; biased progressive sampling between the old tree and a completed subtree; a subtree that turns or
diverges (H - H0 > max_energy_diff = 1000) is discarded and ends the transition; max_tree_depth = 10; the adaptation
statistic is the mean of min(1, exp(H0 - H)) over every leaf evaluated.  Randomness is the sampler's counter-based
Philox stream (csrc/rng.cuh), here evaluated with integer tensor arithmetic, so a transition can be checked draw for
draw against the recursive single-chain oracle.

Everything in this file is tensor bookkeeping on whatever device the state lives on; the arithmetic that costs
anything is inside ``value_and_grad``.  `MAGI_v2.predict(sampler="nuts")` wires that to the CUDA operator only."""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch

Tensor = torch.Tensor

RNG_MOMENTUM, RNG_NUTS_DEPTH, RNG_NUTS_LEAF = 0, 2, 3
_M32 = 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------------------------
# Philox4x32-10 with int64 tensors holding uint32 values (same stream as csrc/rng.cuh)
# ------------------------------------------------------------------------------------------------------------------
def philox4x32_10(c0: Tensor, c1: Tensor, c2: Tensor, c3: Tensor, seed: int):
    """Counters c0..c3: int64 tensors (broadcastable) with values in [0, 2^32).  Returns four int64 tensors."""
    k0, k1 = seed & _M32, (seed >> 32) & _M32
    c0, c1, c2, c3 = torch.broadcast_tensors(c0, c1, c2, c3)
    for _ in range(10):
        # 32 x 32 -> 64-bit products; split the multiplier so that nothing exceeds the signed 64-bit range
        hi0, lo0 = _mulhilo(0xD2511F53, c0)
        hi1, lo1 = _mulhilo(0xCD9E8D57, c2)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + 0x9E3779B9) & _M32
        k1 = (k1 + 0xBB67AE85) & _M32
    return c0, c1, c2, c3


def _mulhilo(a: int, b: Tensor):
    """(high, low) 32-bit halves of a * b for a < 2^32, 0 <= b < 2^32, using signed 64-bit arithmetic."""
    a_hi, a_lo = a >> 16, a & 0xFFFF
    t_lo = a_lo * b                                   # < 2^48
    t_hi = a_hi * b                                   # < 2^48, weight 2^16
    s = t_lo + ((t_hi & 0xFFFF) << 16)                # < 2^49
    lo = s & _M32
    hi = (t_hi >> 16) + (s >> 32)
    return hi, lo


def _u53(hi: Tensor, lo: Tensor) -> Tensor:
    k = (hi << 21) ^ (lo >> 11)
    return (k.to(torch.float64) + 0.5) * (2.0 ** -53)


def rng_normals(seed: int, chain_ids: Tensor, iteration: int, count: int) -> Tensor:
    """[C, count] standard normals: pair j of chain c -> Philox(ctr = (j, c, iteration, 0)) -> Box-Muller."""
    npair = (count + 1) // 2
    dev = chain_ids.device
    j = torch.arange(npair, dtype=torch.int64, device=dev)[None, :]
    it = torch.full((1, 1), int(iteration), dtype=torch.int64, device=dev)
    r = philox4x32_10(j, chain_ids[:, None], it, torch.full_like(it, RNG_MOMENTUM), seed)
    u1, u2 = _u53(r[0], r[1]), _u53(r[2], r[3])
    rad = torch.sqrt(-2.0 * torch.log(u1))
    ang = 2.0 * math.pi * u2
    z = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang)], dim=2).reshape(chain_ids.shape[0], 2 * npair)
    return z[:, :count].contiguous()


def rng_uniform_pairs(seed: int, chain_ids: Tensor, iteration: int, purpose: int, index0: int, count: int):
    """Two [C, count] uniform arrays for counters (index0 + k, chain, iteration, purpose), k < count."""
    dev = chain_ids.device
    k = torch.arange(index0, index0 + count, dtype=torch.int64, device=dev)[None, :]
    it = torch.full((1, 1), int(iteration), dtype=torch.int64, device=dev)
    r = philox4x32_10(k, chain_ids[:, None], it, torch.full_like(it, int(purpose)), seed)
    return _u53(r[0], r[1]), _u53(r[2], r[3])


# ------------------------------------------------------------------------------------------------------------------
# one transition for C chains
# ------------------------------------------------------------------------------------------------------------------
def _dot(a: Tensor, b: Tensor) -> Tensor:
    return torch.einsum("cs,cs->c", a, b)


def _sel(mask: Tensor, new: Tensor, old: Tensor) -> Tensor:
    """In-place old <- where(mask, new, old) for [C, S] (or [C]) tensors; returns old."""
    m = mask if old.dim() == 1 else mask[:, None]
    return old.copy_(torch.where(m, new, old))


def _build_subtree(sub: dict, value_and_grad) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """2^j leaves continuing from (zc, pc, gc) (updated in place) with tensor ops only: the device-agnostic form of
    the loop body (the CUDA form is `FusedLeafEngine.build_subtree`).  Returns (rho_sub, logw_sub, sub_z, sub_lp,
    building): momentum sum, log weight and multinomial proposal of the subtree, and which chains completed it."""
    zc, pc, gc, e, H0, lp0 = sub["zc"], sub["pc"], sub["gc"], sub["e"], sub["H0"], sub["lp0"]
    building, log_u_leaf, n_sub = sub["building"], sub["log_u_leaf"], sub["n_sub"]
    sum_acc, n_leaf, diverged, ck_p, ck_rho = sub["sum_acc"], sub["n_leaf"], sub["diverged"], sub["ck_p"], sub["ck_rho"]
    max_energy_diff, sync_every = sub["max_energy_diff"], sub["sync_every"]
    rho_sub = torch.zeros_like(zc)
    logw_sub = torch.full_like(H0, -math.inf)
    sub_z, sub_lp = zc.clone(), lp0.clone()
    for i in range(n_sub):
        if i % sync_every == 0 and i > 0 and not bool(building.any()):
            break
        even = i % 2 == 0 and n_sub > 1
        if even:
            # an even leaf opens dyadic blocks: remember the momentum sum before it ...
            slot = bin(i).count("1")
            ck_rho[slot].copy_(rho_sub)
        # one leapfrog step (TFP SimpleLeapfrogIntegrator, identity mass) with the signed step size
        ph = pc + (0.5 * e)[:, None] * gc
        zn = zc + e[:, None] * ph
        lpn, gn = value_and_grad(zn)
        pn = ph + (0.5 * e)[:, None] * gn
        if even:
            ck_p[slot].copy_(pn)                             # ... and its own momentum (the block's first)
        dE = -lpn + 0.5 * _dot(pn, pn) - H0
        dE = torch.where(torch.isfinite(dE), dE, torch.full_like(dE, math.inf))
        sum_acc += torch.where(building, torch.exp(torch.clamp(-dE, max=0.0)), torch.zeros_like(dE))
        n_leaf += building.to(torch.int64)
        div = dE > max_energy_diff
        diverged |= building & div
        lw_new = torch.logaddexp(logw_sub, -dE)
        take = building & (log_u_leaf[:, i] < (-dE - lw_new))
        _sel(take, zn, sub_z)
        _sel(take, lpn, sub_lp)
        _sel(building, lw_new, logw_sub)
        rho_sub = rho_sub + torch.where(building[:, None], pn, torch.zeros_like(pn))
        _sel(building, zn, zc); _sel(building, pn, pc); _sel(building, gn, gc)
        ok = ~div
        if i % 2 == 1:
            # dyadic blocks [i - 2^k + 1, i] that end here, k = 1 .. number of trailing ones of i
            t = (~i & (i + 1)).bit_length() - 1
            for k in range(1, t + 1):
                s = i - (1 << k) + 1
                slot = bin(s).count("1")
                rb = rho_sub - ck_rho[slot]
                ok = ok & (_dot(rb, ck_p[slot]) > 0.0) & (_dot(rb, pn) > 0.0)
        building = building & ok
    return rho_sub, logw_sub, sub_z, sub_lp, building


def nuts_transition(z: Tensor, eps: Tensor, value_and_grad: Callable[[Tensor], Tuple[Tensor, Tensor]], seed: int,
                    chain_ids: Tensor, iteration: int, max_tree_depth: int = 10, max_energy_diff: float = 1000.0,
                    sync_every: int = 8, leaf_engine=None, start: Optional[Tuple[Tensor, Tensor]] = None) -> Dict[str, Tensor]:
    """z [C, S] (updated in place to the selected proposals), eps [C] step sizes, value_and_grad(z) -> (lp [C], g [C, S]).
    `start` = (lp0 [C], g0 [C, S]): the target value / gradient at z the transition starts from (TFP keeps them in the
    kernel results of the previous step); None evaluates them here.
    Returns accept_stat [C], n_leapfrog [C] (leaves evaluated while the chain was still building), depth [C],
    diverged [C], lp [C] (log-posterior of the new state)."""
    C, S = z.shape
    dev, f64 = z.device, torch.float64
    lp0, g0 = value_and_grad(z) if start is None else start
    lp0, g0 = lp0.clone(), g0.clone()
    if leaf_engine is None:
        p0 = rng_normals(seed, chain_ids, iteration, S)
    else:
        p0 = leaf_engine.momentum(seed, chain_ids, iteration)
    H0 = -lp0 + 0.5 * _dot(p0, p0)
    uniforms = rng_uniform_pairs if leaf_engine is None else leaf_engine.uniforms
    u_dir, u_acc = uniforms(seed, chain_ids, iteration, RNG_NUTS_DEPTH, 0, max_tree_depth)

    zl, pl, gl = z.clone(), p0.clone(), g0.clone()           # leftmost / rightmost states of the tree
    zr, pr, gr = z.clone(), p0.clone(), g0.clone()
    rho = p0.clone()
    logw = torch.zeros(C, dtype=f64, device=dev)
    prop_z, prop_lp = z.clone(), lp0.clone()
    sum_acc = torch.zeros(C, dtype=f64, device=dev)
    n_leaf = torch.zeros(C, dtype=torch.int64, device=dev)
    depth_out = torch.zeros(C, dtype=torch.int64, device=dev)
    diverged = torch.zeros(C, dtype=torch.bool, device=dev)
    active = torch.ones(C, dtype=torch.bool, device=dev)       # still doubling
    # checkpoint memory: slot popcount(i) holds, for the even leaf i, its momentum and the momentum sum before it
    n_slots = max(max_tree_depth - 1, 1)
    if leaf_engine is None:
        ck_p = torch.empty((n_slots,) + tuple(z.shape), dtype=f64, device=dev)
        ck_rho = torch.empty((n_slots,) + tuple(z.shape), dtype=f64, device=dev)
    else:
        ck_p, ck_rho = leaf_engine.checkpoints(n_slots)

    for j in range(max_tree_depth):
        if not bool(active.any()):
            break
        depth_out += active.to(torch.int64)
        fwd = u_dir[:, j] < 0.5
        e = torch.where(fwd, eps, -eps)
        n_sub = 1 << j
        u_leaf, _ = uniforms(seed, chain_ids, iteration, RNG_NUTS_LEAF, n_sub - 1, n_sub)
        sub = dict(e=e, H0=H0, lp0=lp0, building=active.clone(), log_u_leaf=torch.log(u_leaf),
                   sum_acc=sum_acc, n_leaf=n_leaf, diverged=diverged, ck_p=ck_p, ck_rho=ck_rho, n_sub=n_sub,
                   max_energy_diff=max_energy_diff, sync_every=sync_every)
        log_u_acc = torch.log(u_acc[:, j])
        if leaf_engine is not None:
            # the product path: begin / per-leaf / merge kernels of include/magi_b200_nuts.h
            tree = dict(zl=zl, pl=pl, gl=gl, zr=zr, pr=pr, gr=gr, rho=rho, prop_z=prop_z, prop_lp=prop_lp, logw=logw,
                        fwd=fwd)
            leaf_engine.begin(sub, tree)
            leaf_engine.build_subtree(sub)
            active = leaf_engine.merge(tree, log_u_acc)
            continue
        # the same steps with tensor ops: the subtree starts from the end the direction points to
        fm = fwd[:, None]
        sub["zc"], sub["pc"], sub["gc"] = torch.where(fm, zr, zl), torch.where(fm, pr, pl), torch.where(fm, gr, gl)
        zc, pc, gc = sub["zc"], sub["pc"], sub["gc"]
        rho_sub, logw_sub, sub_z, sub_lp, building = _build_subtree(sub, value_and_grad)
        # merge the completed subtrees
        done = building
        swap = done & (log_u_acc < (logw_sub - logw))
        _sel(swap, sub_z, prop_z)
        _sel(swap, sub_lp, prop_lp)
        _sel(done, torch.logaddexp(logw, logw_sub), logw)
        rho += torch.where(done[:, None], rho_sub, torch.zeros_like(rho_sub))
        mr, ml = done & fwd, done & ~fwd
        _sel(mr, zc, zr); _sel(mr, pc, pr); _sel(mr, gc, gr)
        _sel(ml, zc, zl); _sel(ml, pc, pl); _sel(ml, gc, gl)
        no_turn = (_dot(rho, pl) > 0.0) & (_dot(rho, pr) > 0.0)
        active = done & no_turn
    z.copy_(prop_z)
    acc = sum_acc / torch.clamp(n_leaf, min=1).to(f64)
    return {"accept_stat": acc, "n_leapfrog": n_leaf, "depth": depth_out, "diverged": diverged, "lp": prop_lp}


# ------------------------------------------------------------------------------------------------------------------
# dual averaging (tfp.mcmc.DualAveragingStepSizeAdaptation, per chain) -- same recursion as the fused HMC kernel
# (csrc/sampler.cu) and oracle.dual_averaging_update; da_state [C, 4] = (error_sum, log_averaging_step,
# log_shrinkage_target, step)
# ------------------------------------------------------------------------------------------------------------------
def dual_averaging_update_(eps: Tensor, da: Tensor, accept: Tensor, num_adapt: int, target: float = 0.75,
                           shrinkage: float = 0.05, smoothing: float = 10.0, decay: float = 0.75) -> None:
    step = da[:, 3]
    adapting = step < num_adapt
    err = da[:, 0] + (target - accept)
    t = step + 1.0
    log_x = da[:, 2] - torch.sqrt(t) * err / (shrinkage * (t + smoothing))
    eta = t ** (-decay)
    lavg = eta * log_x + (1.0 - eta) * da[:, 1]
    last = (step + 1.0) == num_adapt
    new_eps = torch.where(last, torch.exp(lavg), torch.exp(log_x))
    da[:, 0] = torch.where(adapting, err, da[:, 0])
    da[:, 1] = torch.where(adapting, lavg, da[:, 1])
    eps.copy_(torch.where(adapting, new_eps, eps))
    da[:, 3] = step + 1.0


def nuts_run_(z: Tensor, eps: Tensor, da: Tensor,
              value_and_grad_at: Optional[Callable[[Tensor, float], Tuple[Tensor, Tensor]]],
              *, n_iter: int, iter0: int = 0, num_adapt: int = 0, min_temp: float = 0.1,
              fixed_beta_temp: Optional[float] = None, target_accept: float = 0.75, seed: int = 0,
              chain_ids: Optional[Tensor] = None, max_tree_depth: int = 10,
              on_sample: Optional[Callable[[int, Tensor, Dict[str, Tensor]], None]] = None,
              leaf_engine: Optional["FusedLeafEngine"] = None, cached_target: bool = False) -> Dict[str, Tensor]:
    """n_iter NUTS transitions in place on z [C, S], eps [C], da [C, 4].  value_and_grad_at(z, beta_temp) evaluates
    the tempered log-posterior (magi_v2.py:348) at temperature beta_temp = max(1 / log(step + 2), min_temp)
    (:833-835, :855) of the global iteration.  `on_sample(it, z, info)` is called after every transition.  With a
    `leaf_engine` (the product path: `FusedLeafEngine`) the evaluations and the per-leaf bookkeeping are the CUDA
    kernels and `value_and_grad_at` is not used.

    cached_target = True reproduces what the reference's wrapper does to TFP's NUTS (magi_v2.py:852-879): the rebuilt
    kernel is handed the PREVIOUS step's kernel results, so a transition starts from the target value and gradient that
    were computed at the previous step's temperature (TFP never re-evaluates them), while every leaf uses the new one.
    Here that is one extra evaluation at the old temperature before the transition (same numbers, no carried state).
    False evaluates the start point at the new temperature."""
    C = z.shape[0]
    if chain_ids is None:
        chain_ids = torch.arange(C, dtype=torch.int64, device=z.device)
    acc = torch.empty((n_iter, C), dtype=torch.float64, device=z.device)
    nleap = torch.empty((n_iter, C), dtype=torch.int64, device=z.device)
    lps = torch.empty((n_iter, C), dtype=torch.float64, device=z.device)
    div = torch.zeros((n_iter, C), dtype=torch.bool, device=z.device)
    for it in range(n_iter):
        g_it = iter0 + it
        bt = float(fixed_beta_temp) if fixed_beta_temp else max(1.0 / math.log(g_it + 2.0), min_temp)
        bt_prev = bt if (fixed_beta_temp or g_it == 0) else max(1.0 / math.log(g_it + 1.0), min_temp)
        stale = cached_target and bt_prev != bt
        if leaf_engine is not None:
            start = None
            if stale:
                leaf_engine.set_beta_temp(bt_prev)
                lp0, g0 = leaf_engine.value_and_grad(z)
                start = (lp0.clone(), g0.clone())
            leaf_engine.set_beta_temp(bt)
            info = nuts_transition(z, eps, leaf_engine.value_and_grad, seed, chain_ids, g_it, max_tree_depth,
                                   leaf_engine=leaf_engine, start=start)
        else:
            start = value_and_grad_at(z, bt_prev) if stale else None
            info = nuts_transition(z, eps, lambda zz: value_and_grad_at(zz, bt), seed, chain_ids, g_it, max_tree_depth,
                                   start=start)
        dual_averaging_update_(eps, da, info["accept_stat"], num_adapt, target_accept)
        acc[it], nleap[it], lps[it], div[it] = info["accept_stat"], info["n_leapfrog"], info["lp"], info["diverged"]
        if on_sample is not None:
            on_sample(it, z, info)
    return {"accept_prob": acc, "n_leapfrog": nleap, "lp": lps, "diverged": div}


# ------------------------------------------------------------------------------------------------------------------
# the product wiring: state = [X | sigma_sqs_pre | thetas_pre] per chain, target = the CUDA operator
# ------------------------------------------------------------------------------------------------------------------
def problem_value_and_grad(prob, R: int) -> Callable[[Tensor, float], Tuple[Tensor, Tensor]]:
    """value_and_grad_at(z [B*R, S], beta_temp) for an `ops.PosteriorProblem`: splits the packed state into the three
    state parts of magi_v2.py:383, calls `magi_b200_logpost_grad` (CUDA; raises without it) and re-packs the gradient."""
    B, n, D, P = prob.B, prob.n, prob.D, prob.P
    nD = n * D

    def value_and_grad_at(z: Tensor, beta_temp: float):
        X = z[:, :nD].reshape(B, R, n, D).contiguous()
        s = z[:, nD:nD + D].reshape(B, R, D).contiguous()
        tau = z[:, nD + D:].reshape(B, R, P).contiguous()
        bt = torch.full((B, R), float(beta_temp), dtype=torch.float64, device=z.device)
        lp, gX, gs, gt = prob.logpost_grad(X, s, tau, bt)
        return lp.reshape(B * R), torch.cat([gX.reshape(B * R, nD), gs.reshape(B * R, D), gt.reshape(B * R, P)], dim=1)

    return value_and_grad_at


def pack_state(X: Tensor, s: Tensor, tau: Tensor) -> Tensor:
    """[B,R,n,D], [B,R,D], [B,R,P] -> z [B*R, n*D + D + P]."""
    B, R = X.shape[:2]
    return torch.cat([X.reshape(B * R, -1), s.reshape(B * R, -1), tau.reshape(B * R, -1)], dim=1).contiguous()


class FusedLeafEngine:
    """The tree builder's steps as CUDA kernels (include/magi_b200_nuts.h): the momentum draw, per doubling
    `magi_b200_nuts_subtree_begin` and `magi_b200_nuts_merge`, and per leaf `magi_b200_nuts_leaf_pre` ->
    `magi_b200_logpost_grad` -> `magi_b200_nuts_leaf_post`.  Every [C, S] array is read or written once per step and
    chains whose subtree has ended are skipped by the bookkeeping kernels.  CUDA only: no CPU fallback."""

    def __init__(self, prob, R: int):
        from . import _lib
        self._lib, self.prob, self.R = _lib, prob, R
        self.C, self.nD, self.D, self.P = prob.B * R, prob.n * prob.D, prob.D, prob.P
        self.S = self.nD + self.D + self.P
        dev, f64 = prob.device, torch.float64
        B, n, D, P = prob.B, prob.n, prob.D, prob.P
        mk = lambda *sh: torch.empty(sh, dtype=f64, device=dev)
        self.ph, self.zc, self.pc, self.gc, self.rho_sub, self.sub_z = (mk(self.C, self.S) for _ in range(6))
        self.sub_lp, self.logw_sub = mk(self.C), mk(self.C)
        self.Xn, self.sn, self.tn = mk(B, R, n, D), mk(B, R, D), mk(B, R, P)
        self.out = prob.logpost_grad_out(R)
        self.bt = mk(B, R)
        self.beta_temp = None
        self._ck = None
        self._st = None
        # per-chain state of the subtree under construction: fixed buffers, one argument struct for every launch
        self.p_sum_acc, self.p_e, self.p_H0 = mk(self.C), mk(self.C), mk(self.C)
        self.p_n_leaf = torch.zeros(self.C, dtype=torch.int64, device=dev)
        self.p_building = torch.zeros(self.C, dtype=torch.bool, device=dev)
        self.p_diverged = torch.zeros(self.C, dtype=torch.bool, device=dev)
        self.p_lu = None                                        # [C, leaves of the largest subtree] log-uniforms
        self.SEG = 8                                            # leaves between early-exit checks

    def _stream(self):
        import ctypes as Ct
        return Ct.c_void_p(torch.cuda.current_stream(self.prob.device).cuda_stream)

    def set_beta_temp(self, beta_temp: float):
        if beta_temp != self.beta_temp:
            self.bt.fill_(float(beta_temp))
            self.beta_temp = beta_temp

    def checkpoints(self, n_slots: int):
        if self._ck is None or self._ck[0].shape[0] < n_slots:
            import ctypes as Ct
            dev = self.prob.device
            self._ck = tuple(torch.empty((n_slots, self.C, self.S), dtype=torch.float64, device=dev) for _ in range(2))
            self.p_lu = torch.zeros((self.C, 1 << n_slots), dtype=torch.float64, device=dev)
            ptr = lambda t: Ct.c_void_p(t.data_ptr())
            self._st = self._lib.NutsSubtree(
                self.C, self.nD, self.D, self.P, ptr(self.zc), ptr(self.pc), ptr(self.gc), ptr(self.rho_sub),
                ptr(self.sub_z), ptr(self.sub_lp), ptr(self.logw_sub), ptr(self.p_sum_acc), ptr(self.p_n_leaf),
                ptr(self.p_building), ptr(self.p_diverged), ptr(self._ck[0]), ptr(self._ck[1]), ptr(self.p_e),
                ptr(self.p_H0))
        return self._ck

    def value_and_grad(self, z: Tensor):
        """Generic evaluation at a packed state (once per transition, for the starting point)."""
        B, R, n, D, P, nD = self.prob.B, self.R, self.prob.n, self.D, self.P, self.nD
        lp, gX, gs, gt = self.prob.logpost_grad(z[:, :nD].reshape(B, R, n, D).contiguous(),
                                                z[:, nD:nD + D].reshape(B, R, D).contiguous(),
                                                z[:, nD + D:].reshape(B, R, P).contiguous(), self.bt)
        return lp.reshape(self.C), torch.cat([gX.reshape(self.C, nD), gs.reshape(self.C, D), gt.reshape(self.C, P)], 1)

    def momentum(self, seed: int, chain_ids: Tensor, iteration: int) -> Tensor:
        import ctypes as Ct
        p0 = torch.empty((self.C, self.S), dtype=torch.float64, device=self.prob.device)
        ids = chain_ids.to(torch.int64).contiguous()
        with torch.cuda.device(self.prob.device):
            self._lib.check(self._lib.lib().magi_b200_nuts_momentum(int(seed), Ct.c_void_p(ids.data_ptr()),
                                                                    int(iteration), self.C, self.S,
                                                                    Ct.c_void_p(p0.data_ptr()), self._stream()),
                            "nuts_momentum")
        return p0

    def uniforms(self, seed: int, chain_ids: Tensor, iteration: int, purpose: int, index0: int, count: int):
        import ctypes as Ct
        ua, ub = (torch.empty((self.C, count), dtype=torch.float64, device=self.prob.device) for _ in range(2))
        ids = chain_ids.to(torch.int64).contiguous()
        with torch.cuda.device(self.prob.device):
            self._lib.check(self._lib.lib().magi_b200_nuts_uniforms(
                int(seed), Ct.c_void_p(ids.data_ptr()), int(iteration), int(purpose), int(index0), int(count), self.C,
                Ct.c_void_p(ua.data_ptr()), Ct.c_void_p(ub.data_ptr()), self._stream()), "nuts_uniforms")
        return ua, ub

    def _tree_struct(self, tree: dict, active: Tensor):
        import ctypes as Ct
        ptr = lambda t: Ct.c_void_p(t.data_ptr())
        for k in ("zl", "pl", "gl", "zr", "pr", "gr", "rho", "prop_z", "prop_lp", "logw", "fwd"):
            if not (tree[k].is_cuda and tree[k].is_contiguous()):
                raise RuntimeError("magi_b200: the fused NUTS path needs contiguous CUDA tensors (no CPU fallback)")
        return self._lib.NutsTree(*(ptr(tree[k]) for k in ("zl", "pl", "gl", "zr", "pr", "gr", "rho", "prop_z",
                                                           "prop_lp", "logw")), ptr(active), ptr(tree["fwd"]))

    def begin(self, sub: dict, tree: dict) -> None:
        import ctypes as Ct
        L = self._lib
        if self._st is None or sub["ck_p"].data_ptr() != self._ck[0].data_ptr():
            raise RuntimeError("magi_b200: the checkpoint memory must come from FusedLeafEngine.checkpoints()")
        n_sub = sub["n_sub"]
        self.p_e.copy_(sub["e"]); self.p_H0.copy_(sub["H0"]); self.p_building.copy_(sub["building"])
        self.p_sum_acc.copy_(sub["sum_acc"]); self.p_n_leaf.copy_(sub["n_leaf"]); self.p_diverged.copy_(sub["diverged"])
        self.p_lu[:, :n_sub].copy_(sub["log_u_leaf"])
        self.sub_lp.copy_(sub["lp0"])
        self.logw_sub.fill_(-math.inf)
        self._active_out = torch.zeros(self.C, dtype=torch.bool, device=self.prob.device)
        tr = self._tree_struct(tree, self._active_out)
        with torch.cuda.device(self.prob.device):
            L.check(L.lib().magi_b200_nuts_subtree_begin(Ct.byref(self._st), Ct.byref(tr), self._stream()),
                    "nuts_subtree_begin")

    def _launch_leaves(self, n_sub: int, i0: int, i1: int, max_energy_diff: float) -> None:
        """Leaves [i0, i1) of a subtree of n_sub leaves on the current stream: three raw C-ABI calls per leaf."""
        import ctypes as Ct
        L, lib = self._lib, self._lib.lib()
        ptr = lambda t: Ct.c_void_p(t.data_ptr())
        lp, gX, gs, gt = self.out
        evalf, ws, nb = self.prob.eval_call(self.R)             # few datasets: the wide path (magi_b200_wide.h)
        pb = self.prob.struct(self.R)
        st_ref, pb_ref = Ct.byref(self._st), Ct.byref(pb)
        p_ph, p_Xn, p_sn, p_tn, p_bt = ptr(self.ph), ptr(self.Xn), ptr(self.sn), ptr(self.tn), ptr(self.bt)
        p_lp, p_gX, p_gs, p_gt, p_ws = ptr(lp), ptr(gX), ptr(gs), ptr(gt), (ptr(ws) if ws is not None else None)
        lu0, lu_stride = self.p_lu.data_ptr(), self.p_lu.shape[1]
        pre, post = lib.magi_b200_nuts_leaf_pre, lib.magi_b200_nuts_leaf_post_next
        no_slots = (Ct.c_int * 1)(0)
        stream = self._stream()
        for i in range(i0, i1):
            # the first half of the step: a launch of its own for the first leaf, fused into the previous leaf's
            # `leaf_post` for the others
            rc = pre(st_ref, p_ph, p_Xn, p_sn, p_tn, stream) if i == 0 else 0
            rc = rc or evalf(pb_ref, p_Xn, p_sn, p_tn, p_bt, p_lp, p_gX, p_gs, p_gt, p_ws, nb, stream)
            if i & 1:
                t = (~i & (i + 1)).bit_length() - 1
                slots = [bin(i - (1 << k) + 1).count("1") for k in range(1, t + 1)]
                arr, slot_store = (Ct.c_int * t)(*slots), -1
            else:
                t, arr = 0, no_slots
                slot_store = bin(i).count("1") if n_sub > 1 else -1
            rc = rc or post(st_ref, p_ph, p_Xn, p_sn, p_tn, p_lp, p_gX, p_gs, p_gt, Ct.c_void_p(lu0 + 8 * i),
                            lu_stride, float(max_energy_diff), slot_store, t, arr, 1 if i + 1 < n_sub else 0, stream)
            if rc:
                L.check(rc, "nuts leaf (leaf_pre / logpost_grad / leaf_post)")

    def build_subtree(self, sub: dict) -> None:
        """The leaves of one doubling, in segments of SEG leaves; between segments the host checks whether any chain
        is still building.  (Replaying the segments as CUDA graphs was measured and dropped: for one dataset a leaf is
        bound by the six dependent kernels on the device, 47 us, not by their launches -- profiles/r01_notes.md.)"""
        n_sub, med = sub["n_sub"], float(sub["max_energy_diff"])
        with torch.cuda.device(self.prob.device):
            for i0 in range(0, n_sub, self.SEG):
                if i0 > 0 and not bool(self.p_building.any()):
                    break
                self._launch_leaves(n_sub, i0, min(n_sub, i0 + self.SEG), med)
        sub["sum_acc"].copy_(self.p_sum_acc); sub["n_leaf"].copy_(self.p_n_leaf); sub["diverged"].copy_(self.p_diverged)
        sub["building"].copy_(self.p_building)

    def merge(self, tree: dict, log_u_acc: Tensor) -> Tensor:
        import ctypes as Ct
        L = self._lib
        tr = self._tree_struct(tree, self._active_out)
        lu = log_u_acc.contiguous()
        with torch.cuda.device(self.prob.device):
            L.check(L.lib().magi_b200_nuts_merge(Ct.byref(self._st), Ct.byref(tr), Ct.c_void_p(lu.data_ptr()),
                                                 self._stream()), "nuts_merge")
        return self._active_out
