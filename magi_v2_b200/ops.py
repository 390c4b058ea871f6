"""The C-ABI entry points exposed as PyTorch custom operators, ``torch.ops.magi_b200.*``.

PyTorch is plumbing here (device memory, streams); each op validates its tensors, then passes raw
device pointers and the current CUDA stream to libmagi_b200.so.  Ops exist for CUDA tensors only
(no CPU kernels are registered: calling one with CPU tensors raises NotImplementedError)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import HmcConfig, Problem, check, lib

Tensor = torch.Tensor


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk(t: Tensor, name: str, dtype=torch.float64, shape=None):
    if not t.is_cuda:
        raise RuntimeError(f"magi_b200: {name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"magi_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"magi_b200: {name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"magi_b200: {name} must have shape {tuple(shape)}, got {tuple(t.shape)}")


# ------------------------------------------------------------------------------------------------
# (1) covariance build
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("magi_b200::cov_build", mutates_args=(), device_types="cuda")
def cov_build(I: Tensor, phi1: Tensor, phi2: Tensor, nu: float, uniform_grid: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """I [n] (shared grid) or [B,n]; phi1, phi2 [B,D] -> C, Cp, Cpp each [B,D,n,n]."""
    _chk(phi1, "phi1"); _chk(phi2, "phi2", shape=phi1.shape); _chk(I, "I")
    B, D = phi1.shape
    n = I.shape[-1]
    stride = 0 if I.dim() == 1 else n
    if I.dim() == 2 and I.shape[0] != B:
        raise RuntimeError("magi_b200: I must be [n] or [B,n]")
    with torch.cuda.device(phi1.device):
        out = [torch.empty((B, D, n, n), dtype=torch.float64, device=phi1.device) for _ in range(3)]
        st = lib().magi_b200_cov_build(_ptr(I), stride, _ptr(phi1), _ptr(phi2), float(nu), B, D, n,
                                       _lib.COV_UNIFORM_GRID if uniform_grid else 0,
                                       _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream(phi1))
    check(st, "cov_build")
    return out[0], out[1], out[2]


@cov_build.register_fake
def _(I, phi1, phi2, nu, uniform_grid):
    B, D = phi1.shape
    n = I.shape[-1]
    return tuple(phi1.new_empty((B, D, n, n)) for _ in range(3))


# ------------------------------------------------------------------------------------------------
# (2) factorise + derive
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("magi_b200::factor_derive", mutates_args=(), device_types="cuda")
def factor_derive(C_: Tensor, Cp: Tensor, Cpp: Tensor, band: int, jitter: float) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """C, Cp, Cpp [...,n,n] -> (Cinv, m, Kinv, K, info[...]) ; band < 0 = no banding."""
    _chk(C_, "C"); _chk(Cp, "Cp", shape=C_.shape); _chk(Cpp, "Cpp", shape=C_.shape)
    n = C_.shape[-1]
    nmat = C_.numel() // (n * n)
    with torch.cuda.device(C_.device):
        Cinv, m, Kinv, K = (torch.empty_like(C_) for _ in range(4))
        info = torch.empty(C_.shape[:-2], dtype=torch.int32, device=C_.device)
        wsb = lib().magi_b200_factor_workspace_bytes(nmat, n)
        ws = torch.empty(wsb // 8, dtype=torch.float64, device=C_.device)
        st = lib().magi_b200_factor_derive(_ptr(C_), _ptr(Cp), _ptr(Cpp), nmat, n, int(band), float(jitter),
                                           _ptr(Cinv), _ptr(m), _ptr(Kinv), _ptr(K), _ptr(info), _ptr(ws), wsb,
                                           _stream(C_))
    check(st, "factor_derive")
    return Cinv, m, Kinv, K, info


def spd_inverse(A: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """A [...,n,n] symmetric positive definite -> (A^-1 [...,n,n], log det A [...], info [...]) by the library's blocked
    Cholesky (`magi_b200_spd_inverse`; info > 0: not positive definite at that pivot)."""
    _chk(A, "A")
    n = A.shape[-1]
    nmat = A.numel() // (n * n)
    with torch.cuda.device(A.device):
        Ainv = torch.empty_like(A)
        logdet = torch.empty(A.shape[:-2], dtype=torch.float64, device=A.device)
        info = torch.empty(A.shape[:-2], dtype=torch.int32, device=A.device)
        wsb = lib().magi_b200_factor_workspace_bytes(nmat, n)
        ws = torch.empty(wsb // 8, dtype=torch.float64, device=A.device)
        st = lib().magi_b200_spd_inverse(_ptr(A), nmat, n, _ptr(Ainv), _ptr(logdet), _ptr(info), _ptr(ws), wsb, _stream(A))
    check(st, "spd_inverse")
    return Ainv, logdet, info


@factor_derive.register_fake
def _(C_, Cp, Cpp, band, jitter):
    return (torch.empty_like(C_), torch.empty_like(C_), torch.empty_like(C_), torch.empty_like(C_),
            C_.new_empty(C_.shape[:-2], dtype=torch.int32))


# ------------------------------------------------------------------------------------------------
# (3a) pack
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("magi_b200::pack_matrices", mutates_args=(), device_types="cuda")
def pack_matrices(Cinv: Tensor, m: Tensor, Kinv: Tensor) -> Tensor:
    """Cinv, m, Kinv [B,D,n,n] -> opaque packed buffer (float64, 1-D)."""
    _chk(Cinv, "Cinv"); _chk(m, "m", shape=Cinv.shape); _chk(Kinv, "Kinv", shape=Cinv.shape)
    if Cinv.dim() != 4:
        raise RuntimeError("magi_b200: matrices must be [B,D,n,n]")
    B, D, n, _ = Cinv.shape
    with torch.cuda.device(Cinv.device):
        packed = torch.empty(lib().magi_b200_packed_bytes(B, D, n) // 8, dtype=torch.float64, device=Cinv.device)
        st = lib().magi_b200_pack_matrices(_ptr(Cinv), _ptr(m), _ptr(Kinv), B, D, n, _ptr(packed), _stream(Cinv))
    check(st, "pack_matrices")
    return packed


@pack_matrices.register_fake
def _(Cinv, m, Kinv):
    B, D, n, _ = Cinv.shape
    npad = (n + 7) // 8 * 8
    return Cinv.new_empty((B * D * 3 * npad * npad,))


# ------------------------------------------------------------------------------------------------
# problem constants
# ------------------------------------------------------------------------------------------------
class PosteriorProblem:
    """Device-resident constants of the log-posterior for B datasets (what the reference's
    ``unnormalized_log_prob`` closes over, magi_v2.py:294-300) + the ctypes view of them."""

    def __init__(self, model, packed: Tensor, mu: Tensor, y: Tensor, mask: Tensor, N_ds: Tensor,
                 beta: Tensor, LB: Tensor, n: int, band: Optional[int] = None):
        """`model`: the name of a compiled-in system, or a models.OdeModel -- for a user system (tracing.py) the
        evaluation runs in the library compiled for it (wide path only)."""
        self._ulib = None
        if isinstance(model, str):
            self.model = model
            self.model_id = _lib.MODEL_IDS[model]
            D_, P_ = C.c_int(), C.c_int()
            lib().magi_b200_model_dims(self.model_id, C.byref(D_), C.byref(P_))
            self.D, self.P = D_.value, P_.value
        else:
            self.model, self.model_id, self.D, self.P = model.name, int(model.model_id), int(model.D), int(model.P)
            if model.lib_path is not None:
                self._ulib = _lib.load_user_library(model.lib_path)
        self.B, self.n = mu.shape[0], int(n)
        _chk(mu, "mu", shape=(self.B, self.D)); _chk(y, "y", shape=(self.B, self.n, self.D))
        _chk(mask, "mask", dtype=torch.uint8, shape=(self.B, self.n, self.D))
        _chk(N_ds, "N_ds", shape=(self.B, self.D)); _chk(beta, "beta", shape=(self.B,))
        _chk(LB, "LB", shape=(self.B, self.D)); _chk(packed, "packed")
        if packed.numel() * 8 != lib().magi_b200_packed_bytes(self.B, self.D, self.n):
            raise RuntimeError("magi_b200: packed buffer has the wrong size for (B, D, n)")
        self.packed, self.mu, self.y, self.mask, self.N_ds, self.beta, self.LB = packed, mu, y, mask, N_ds, beta, LB
        self.device = mu.device
        # bandsize the matrices were banded with (None = dense): lets the kernels skip all-zero tiles.
        # The caller vouches for it -- entries outside the band are never read.
        self.band = -1 if band is None else int(band)
        self._ws = {}
        self._host = None

    def set_LB(self, LB) -> None:
        """Replace sigma_sqs_LB (magi_v2.py:299-300: it depends on the smoothed Xhat_init, known last)."""
        LB = torch.as_tensor(LB, dtype=torch.float64, device=self.device).contiguous()
        _chk(LB, "LB", shape=(self.B, self.D))
        self.LB.copy_(LB)

    def struct(self, R: int, b0: int = 0, b1: Optional[int] = None) -> Problem:
        """magi_problem_t for datasets [b0, b1) (a view: pointers offset into the same buffers)."""
        b1 = self.B if b1 is None else b1
        npad = (self.n + 7) // 8 * 8
        D, n = self.D, self.n
        return Problem(self.model_id, b1 - b0, int(R), n, D, self.P,
                       self.packed.data_ptr() + b0 * D * 3 * npad * npad * 8,
                       self.mu.data_ptr() + b0 * D * 8, self.y.data_ptr() + b0 * n * D * 8,
                       self.mask.data_ptr() + b0 * n * D, self.N_ds.data_ptr() + b0 * D * 8,
                       self.beta.data_ptr() + b0 * 8, self.LB.data_ptr() + b0 * D * 8, self.band)

    def workspace(self, R: int, slot: int = 0, n_datasets: Optional[int] = None) -> Tuple[Optional[Tensor], int]:
        """Caller-owned scratch of the sampler kernels; one buffer per concurrent stream (`slot`)."""
        pb = self.struct(R, 0, n_datasets)
        nbytes = lib().magi_b200_sampler_workspace_bytes(C.byref(pb))
        ws = self._ws.get(slot)
        if ws is None or ws.numel() * 8 < nbytes:
            ws = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=self.device)
            self._ws[slot] = ws
        return ws, nbytes

    # -- (3b) ------------------------------------------------------------------------------------
    def logpost_grad_out(self, R: int):
        """Preallocated device outputs (lp [B,R], gX [B,R,n,D], gsig [B,R,D], gth [B,R,P])."""
        mk = lambda *sh: torch.empty(sh, dtype=torch.float64, device=self.device)
        return mk(self.B, R), mk(self.B, R, self.n, self.D), mk(self.B, R, self.D), mk(self.B, R, self.P)

    def eval_path(self, R: int, path: str = "auto") -> str:
        """"cta": magi_b200_logpost_grad, one CTA per (dataset, 8 chains) -- for many datasets; "wide":
        magi_b200_logpost_grad_wide (include/magi_b200_wide.h), every component's matrix rows spread over the grid
        -- for few.  Both compute the same function; "auto" picks by how many CTAs the first one would have.  The
        rule is the measured cross-over (tools/time_wide.py, tools/path_ab.py, profiles/r01_notes.md): at n = 161 one
        CTA streams its dataset in ~75 us whatever B is and the wide path needs ~35 + B us, while with many datasets
        the register-resident fast kernel of the CTA path is ~15 % ahead."""
        if self._ulib is not None:
            if path == "cta":
                raise ValueError("a user-supplied ODE system runs on the wide path only")
            return "wide"
        if path == "auto":
            ctas = self.B * ((R + 7) // 8)
            npad = (self.n + 7) // 8 * 8
            # np > 168: the one-CTA-per-dataset kernels keep their vectors in a global workspace (general path) and
            # the wide path is faster at every batch size measured (SIRW n = 321, B = 512: 1.49 vs 2.24 ms)
            return "wide" if (npad > 168 or ctas <= 32) else "cta"
        if path not in ("cta", "wide"):
            raise ValueError("path must be 'auto', 'cta' or 'wide'")
        return path

    def eval_call(self, R: int, path: str = "auto"):
        """(C entry point, workspace tensor or None, workspace bytes) of the evaluation path for R chains."""
        if self.eval_path(R, path) == "cta":
            ws, nb = self.workspace(R)
            return lib().magi_b200_logpost_grad, ws, nb
        pb = self.struct(R)
        L = self._ulib if self._ulib is not None else lib()
        nb = L.magi_b200_logpost_grad_wide_workspace_bytes(C.byref(pb))
        ws = self._ws.get("wide")
        if ws is None or ws.numel() * 8 < nb:
            ws = torch.empty(max(nb // 8, 1), dtype=torch.float64, device=self.device)
            self._ws["wide"] = ws
        return L.magi_b200_logpost_grad_wide, ws, nb

    def logpost_grad(self, X: Tensor, sig_pre: Tensor, th_pre: Tensor, beta_temp: Tensor, out=None, path: str = "auto"):
        """X [B,R,n,D], sig_pre [B,R,D], th_pre [B,R,P], beta_temp [B,R] ->
        (lp [B,R], gX, gsig, gth) -- value and gradient of magi_v2.py:308-348."""
        R = X.shape[1]
        _chk(X, "X", shape=(self.B, R, self.n, self.D)); _chk(sig_pre, "sig_pre", shape=(self.B, R, self.D))
        _chk(th_pre, "th_pre", shape=(self.B, R, self.P)); _chk(beta_temp, "beta_temp", shape=(self.B, R))
        with torch.cuda.device(self.device):
            lp, gX, gsig, gth = self.logpost_grad_out(R) if out is None else out
            if out is not None:
                _chk(lp, "lp", shape=(self.B, R)); _chk(gX, "gX", shape=X.shape)
                _chk(gsig, "gsig", shape=sig_pre.shape); _chk(gth, "gth", shape=th_pre.shape)
            fn, ws, nb = self.eval_call(R, path)
            pb = self.struct(R)
            st = fn(C.byref(pb), _ptr(X), _ptr(sig_pre), _ptr(th_pre), _ptr(beta_temp),
                    _ptr(lp), _ptr(gX), _ptr(gsig), _ptr(gth), _ptr(ws), nb, _stream(X))
        check(st, "logpost_grad")
        return lp, gX, gsig, gth

    # -- (3b) with HOST buffers: what a host-side sampler (the reference's TFP loop) would call ----------
    def host_pipeline(self, R: int, n_chunks: int = 16, n_streams: int = 4) -> "HostPipeline":
        """Pinned host staging + device mirrors for `logpost_grad` on HOST data (see HostPipeline)."""
        key = (int(R), int(n_chunks), int(n_streams))
        if self._host is None or self._host.key != key:
            self._host = HostPipeline(self, *key)
        return self._host

    def logpost_grad_host_out(self, R: int):
        """Host outputs for `logpost_grad_host`."""
        mk = lambda *sh: torch.empty(sh, dtype=torch.float64)
        return mk(self.B, R), mk(self.B, R, self.n, self.D), mk(self.B, R, self.D), mk(self.B, R, self.P)

    def logpost_grad_host(self, X: Tensor, sig_pre: Tensor, th_pre: Tensor, beta_temp: Tensor, out=None,
                          n_chunks: int = 16, n_streams: int = 4):
        """`logpost_grad` for ordinary HOST tensors: copies them into the pipeline's pinned staging blocks, runs it and
        copies the results out.  Convenience form -- a caller that wants the full PCIe rate fills
        `host_pipeline(R).inputs(c)` in place and reads `.outputs(c)` (no host-side copies).  Returns host tensors
        (lp, gX, gsig, gth); they are complete when the call returns."""
        R = X.shape[1]
        for t, nm, shp in ((X, "X", (self.B, R, self.n, self.D)), (sig_pre, "sig_pre", (self.B, R, self.D)),
                           (th_pre, "th_pre", (self.B, R, self.P)), (beta_temp, "beta_temp", (self.B, R))):
            if t.is_cuda or t.dtype != torch.float64 or tuple(t.shape) != shp:
                raise RuntimeError(f"magi_b200: {nm} must be a float64 host tensor of shape {shp}")
        hp = self.host_pipeline(R, n_chunks, n_streams)
        hp.fill(X, sig_pre, th_pre, beta_temp)
        hp.run()
        return hp.gather(out)

    # -- (3c) ------------------------------------------------------------------------------------
    def leapfrog_(self, X, sig_pre, th_pre, pX, psig, pth, eps, beta_temp, n_steps: int):
        """In-place leapfrog trajectory with the given momenta; returns lp at the end point."""
        if self._ulib is not None:
            raise NotImplementedError("the fused leapfrog kernel exists for the compiled-in systems; a user-supplied "
                                      "system samples through hmc_run_ / nuts (one evaluation launch per step)")
        R = X.shape[1]
        for t, nm, shp in ((X, "X", (self.B, R, self.n, self.D)), (pX, "pX", (self.B, R, self.n, self.D)),
                           (sig_pre, "sig_pre", (self.B, R, self.D)), (psig, "psig", (self.B, R, self.D)),
                           (th_pre, "th_pre", (self.B, R, self.P)), (pth, "pth", (self.B, R, self.P)),
                           (eps, "eps", (self.B, R)), (beta_temp, "beta_temp", (self.B, R))):
            _chk(t, nm, shape=shp)
        with torch.cuda.device(self.device):
            lp = torch.empty((self.B, R), dtype=torch.float64, device=self.device)
            ws, nb = self.workspace(R)
            pb = self.struct(R)
            st = lib().magi_b200_leapfrog(C.byref(pb), _ptr(X), _ptr(sig_pre), _ptr(th_pre), _ptr(pX), _ptr(psig),
                                          _ptr(pth), _ptr(eps), _ptr(beta_temp), int(n_steps), _ptr(lp), _ptr(ws),
                                          nb, _stream(X))
        check(st, "leapfrog")
        return lp

    # -- (3d) ------------------------------------------------------------------------------------
    def hmc_run_(self, X, sig_pre, th_pre, eps, da_state, *, n_iter: int, n_leapfrog: int, iter0: int = 0,
                 num_adapt: int = 0, accum_from: int = 0, min_temp: float = 0.1, fixed_beta_temp: float = 0.0,
                 target_accept: float = 0.75, seed: int = 0, chain_id0: int = 0, keep_theta=True,
                 keep_sigma=True, keep_X=False, X_sum: Optional[Tensor] = None, X_sumsq: Optional[Tensor] = None,
                 path: str = "auto"):
        """Run n_iter HMC transitions in place on (X, sig_pre, th_pre, eps, da_state); returns a dict of
        traces (thetas_samps, sigma_sqs_samps, X_samps, accept_prob, lp).  path "cta": the whole chain inside the
        fused kernel `magi_b200_hmc_run`; "wide": one launch of the wide evaluation per leapfrog step
        (hmc_host.py) -- same algorithm, same draws.  "auto" takes the second only for np > 168, where the wide
        evaluation is several times faster than the fused kernel's general path; for small grids the fused kernel
        wins even for one dataset because a whole chain costs one launch."""
        R = X.shape[1]
        if path not in ("auto", "cta", "wide"):
            raise ValueError("path must be 'auto', 'cta' or 'wide'")
        if self._ulib is not None or path == "wide" or (path == "auto" and (self.n + 7) // 8 * 8 > 168):
            from .hmc_host import hmc_run_host_
            return hmc_run_host_(self, X, sig_pre, th_pre, eps, da_state, n_iter=n_iter, n_leapfrog=n_leapfrog,
                                 iter0=iter0, num_adapt=num_adapt, accum_from=accum_from, min_temp=min_temp,
                                 fixed_beta_temp=fixed_beta_temp, target_accept=target_accept, seed=seed,
                                 chain_id0=chain_id0, keep_theta=keep_theta, keep_sigma=keep_sigma, keep_X=keep_X,
                                 X_sum=X_sum, X_sumsq=X_sumsq)
        _chk(X, "X", shape=(self.B, R, self.n, self.D)); _chk(sig_pre, "sig_pre", shape=(self.B, R, self.D))
        _chk(th_pre, "th_pre", shape=(self.B, R, self.P)); _chk(eps, "eps", shape=(self.B, R))
        _chk(da_state, "da_state", shape=(self.B, R, 4))
        dev = self.device
        with torch.cuda.device(dev):
            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
            th_s = mk(n_iter, self.B, R, self.P) if keep_theta else None
            sg_s = mk(n_iter, self.B, R, self.D) if keep_sigma else None
            X_s = mk(n_iter, self.B, R, self.n, self.D) if keep_X else None
            acc, lpt = mk(n_iter, self.B, R), mk(n_iter, self.B, R)
            cfg = HmcConfig(int(n_iter), int(n_leapfrog), int(iter0), int(num_adapt), int(accum_from),
                            float(min_temp), float(fixed_beta_temp), float(target_accept), int(seed),
                            int(chain_id0))
            ws, nb = self.workspace(R)
            pb = self.struct(R)
            st = lib().magi_b200_hmc_run(C.byref(pb), C.byref(cfg), _ptr(X), _ptr(sig_pre), _ptr(th_pre), _ptr(eps),
                                         _ptr(da_state), _ptr(th_s), _ptr(sg_s), _ptr(X_s), _ptr(X_sum),
                                         _ptr(X_sumsq), _ptr(acc), _ptr(lpt), _ptr(ws), nb, _stream(X))
        check(st, "hmc_run")
        return {"thetas_samps": th_s, "sigma_sqs_samps": sg_s, "X_samps": X_s, "accept_prob": acc, "lp": lpt}


# ------------------------------------------------------------------------------------------------
# (3b) as a registered operator (functional form; the class above is the stateful convenience)
# ------------------------------------------------------------------------------------------------
class HostPipeline:
    """`magi_b200_logpost_grad` for chain states that live in HOST memory (the position of a host-side sampler such as
    the reference's TFP loop, which calls its target once per leapfrog step, magi_v2.py:362).

    The batch is cut into `n_chunks` dataset chunks.  Every chunk owns ONE contiguous pinned input block
    [X | sig_pre | th_pre | beta_temp] and ONE contiguous pinned output block [lp | gX | gsig | gth], mirrored on the
    device, so that a chunk costs exactly one host->device copy, one kernel launch and one device->host copy;
    chunks are pipelined over `n_streams` CUDA streams (both PCIe directions and the kernels overlap).
    `inputs(c)` / `outputs(c)` are views into the pinned blocks: fill / read them in place.  `run()` returns when
    every output block is complete on the host."""

    def __init__(self, prob: "PosteriorProblem", R: int, n_chunks: int, n_streams: int, ws_slot0: int = 1):
        # ws_slot0: first workspace slot of this pipeline (two pipelines that run concurrently -- double buffering:
        # one uploads / evaluates batch k+1 while the other still downloads batch k -- must not share scratch)
        self.prob, self.R, self.key, self.ws_slot0 = prob, R, (R, n_chunks, n_streams), ws_slot0
        B, n, D, P = prob.B, prob.n, prob.D, prob.P
        n_chunks = max(1, min(n_chunks, B))
        per = -(-B // n_chunks)
        self.bounds = [(b0, min(B, b0 + per)) for b0 in range(0, B, per)]
        self.shapes = lambda nb: ((nb, R, n, D), (nb, R, D), (nb, R, P), (nb, R))
        self.oshapes = lambda nb: ((nb, R), (nb, R, n, D), (nb, R, D), (nb, R, P))
        per_chain = n * D + D + P + 1
        self.offsets = [b0 * R * per_chain for b0, _ in self.bounds] + [B * R * per_chain]
        total = B * R * per_chain
        self.h_in = torch.empty(total, dtype=torch.float64).pin_memory()
        self.h_out = torch.empty(total, dtype=torch.float64).pin_memory()
        with torch.cuda.device(prob.device):
            self.d_in = torch.empty(total, dtype=torch.float64, device=prob.device)
            self.d_out = torch.empty(total, dtype=torch.float64, device=prob.device)
            self.streams = [torch.cuda.Stream(prob.device) for _ in range(max(1, n_streams))]
        self.done = torch.cuda.Event()
        self.h2d_bytes = self.d2h_bytes = total * 8

    @staticmethod
    def _views(flat: Tensor, off: int, shapes):
        out = []
        for sh in shapes:
            k = 1
            for v in sh:
                k *= v
            out.append(flat[off:off + k].view(sh))
            off += k
        return tuple(out)

    @property
    def n_chunks(self) -> int:
        return len(self.bounds)

    def inputs(self, c: int):
        """(X, sig_pre, th_pre, beta_temp) of chunk c: pinned host views, datasets bounds[c][0] .. bounds[c][1]."""
        b0, b1 = self.bounds[c]
        return self._views(self.h_in, self.offsets[c], self.shapes(b1 - b0))

    def outputs(self, c: int):
        """(lp, gX, gsig, gth) of chunk c: pinned host views."""
        b0, b1 = self.bounds[c]
        return self._views(self.h_out, self.offsets[c], self.oshapes(b1 - b0))

    def fill(self, X: Tensor, sig_pre: Tensor, th_pre: Tensor, beta_temp: Tensor) -> None:
        for c, (b0, b1) in enumerate(self.bounds):
            for dst, src in zip(self.inputs(c), (X, sig_pre, th_pre, beta_temp)):
                dst.copy_(src[b0:b1])

    def gather(self, out=None):
        out = self.prob.logpost_grad_host_out(self.R) if out is None else out
        for c, (b0, b1) in enumerate(self.bounds):
            for dst, src in zip(out, self.outputs(c)):
                dst[b0:b1].copy_(src)
        return out

    def run(self, wait: bool = True) -> None:
        prob, R = self.prob, self.R
        with torch.cuda.device(prob.device):
            cur = torch.cuda.current_stream(prob.device)
            start = torch.cuda.Event()
            start.record(cur)
            per = self.bounds[0][1] - self.bounds[0][0]
            for c, (b0, b1) in enumerate(self.bounds):
                stq = self.streams[c % len(self.streams)]
                stq.wait_event(start)
                o0, o1 = self.offsets[c], self.offsets[c + 1]
                with torch.cuda.stream(stq):
                    self.d_in[o0:o1].copy_(self.h_in[o0:o1], non_blocking=True)           # one H2D
                    dX, ds, dt, dbt = self._views(self.d_in, o0, self.shapes(b1 - b0))
                    lp, gX, gs, gt = self._views(self.d_out, o0, self.oshapes(b1 - b0))
                    ws, nb = prob.workspace(R, slot=self.ws_slot0 + c % len(self.streams), n_datasets=per)
                    pb = prob.struct(R, b0, b1)
                    st = lib().magi_b200_logpost_grad(C.byref(pb), _ptr(dX), _ptr(ds), _ptr(dt), _ptr(dbt), _ptr(lp),
                                                      _ptr(gX), _ptr(gs), _ptr(gt), _ptr(ws), nb,
                                                      C.c_void_p(stq.cuda_stream))
                    check(st, "logpost_grad")
                    self.h_out[o0:o1].copy_(self.d_out[o0:o1], non_blocking=True)         # one D2H
            for stq in self.streams:
                ev = torch.cuda.Event()
                ev.record(stq)
                cur.wait_event(ev)
            self.done.record(cur)
        if wait:
            self.done.synchronize()


@torch.library.custom_op("magi_b200::logpost_grad", mutates_args=(), device_types="cuda")
def logpost_grad(model_id: int, X: Tensor, sig_pre: Tensor, th_pre: Tensor, beta_temp: Tensor, packed: Tensor,
                 mu: Tensor, y: Tensor, mask: Tensor, N_ds: Tensor, beta: Tensor, LB: Tensor, band: int = -1) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    name = {v: k for k, v in _lib.MODEL_IDS.items()}[model_id]
    prob = PosteriorProblem(name, packed, mu, y, mask, N_ds, beta, LB, X.shape[2], None if band < 0 else band)
    return prob.logpost_grad(X, sig_pre, th_pre, beta_temp)


@logpost_grad.register_fake
def _(model_id, X, sig_pre, th_pre, beta_temp, packed, mu, y, mask, N_ds, beta, LB, band=-1):
    return X.new_empty(X.shape[:2]), torch.empty_like(X), torch.empty_like(sig_pre), torch.empty_like(th_pre)
