"""ctypes binding of oracle/magi_oracle_c.c (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

``build()`` compiles it with the Makefile beside it (gcc only); ``COracle(constants, model)`` wraps one
dataset's posterior: ``logpost_grad``, ``leapfrog``, ``nuts_chain``, ``hmc_chain``."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import magi_oracle as mo

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmagi_oracle_c.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "magi_oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.mo_create.restype = C.c_void_p
        L.mo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                C.POINTER(C.c_ubyte), _dp, C.c_double, _dp]
        L.mo_destroy.argtypes = [C.c_void_p]
        L.mo_state_size.argtypes = [C.c_void_p]
        L.mo_logpost_grad.restype = C.c_double
        L.mo_logpost_grad.argtypes = [C.c_void_p, _dp, C.c_double, _dp]
        L.mo_logpost_grad_batch.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
        L.mo_rng_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _dp]
        L.mo_leapfrog.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.c_double, _dp]
        L.mo_nuts_chain.argtypes = [C.c_void_p, _dp, C.c_int, C.c_double, C.c_uint64, C.c_uint32, C.c_int, C.c_double,
                                    C.c_int, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip, _ip, _dp]
        L.mo_hmc_chain.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_uint32, C.c_int,
                                   C.c_double, C.c_int, C.c_double, _dp, _dp, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class COracle:
    """One dataset's posterior in the C oracle.  ``c``: oracle PosteriorConstants; ``band``: the bandsize the
    matrices were cut to (lets the loops skip the zeros; None = dense)."""

    def __init__(self, c: mo.PosteriorConstants, model: str, band=None):
        self.L = lib()
        self.model = mo.MODELS[model]
        self.n, self.D, self.P = c.n, c.D, self.model.P
        y, mask = c.dense_y_mask()
        args = [_f64(c.C_d_invs), _f64(c.m_ds), _f64(c.K_d_invs), _f64(c.mu_ds), _f64(y),
                np.ascontiguousarray(mask, dtype=np.uint8), _f64(c.N_ds), _f64(c.sigma_sqs_LB)]
        self._keep = args
        self.h = self.L.mo_create(self.n, self.D, self.P, self.model.model_id, -1 if band is None else int(band),
                                  _d(args[0]), _d(args[1]), _d(args[2]), _d(args[3]), _d(args[4]),
                                  args[5].ctypes.data_as(C.POINTER(C.c_ubyte)), _d(args[6]), float(c.beta), _d(args[7]))
        if not self.h:
            raise RuntimeError("mo_create failed")
        self.S = self.L.mo_state_size(self.h)

    def __del__(self):
        try:
            self.L.mo_destroy(self.h)
        except Exception:
            pass

    def logpost_grad(self, z, beta_temp):
        z = _f64(z)
        g = np.empty(self.S)
        lp = self.L.mo_logpost_grad(self.h, _d(z), float(beta_temp), _d(g))
        return lp, g

    def logpost_grad_batch(self, Z, beta_temp):
        Z = _f64(Z)
        R = Z.shape[0]
        bt = _f64(np.broadcast_to(beta_temp, (R,)))
        lp, G = np.empty(R), np.empty((R, self.S))
        self.L.mo_logpost_grad_batch(self.h, R, _d(Z), _d(bt), _d(lp), _d(G))
        return lp, G

    def leapfrog(self, z, p, eps, n_steps, beta_temp):
        z, p = _f64(z).copy(), _f64(p).copy()
        traj = np.empty((n_steps, self.S))
        self.L.mo_leapfrog(self.h, _d(z), _d(p), float(eps), int(n_steps), float(beta_temp), _d(traj))
        return z, p, traj

    def nuts_chain(self, z0, n_iter, eps0=0.1, seed=0, chain_id=0, num_adaptation_steps=0, min_temp=0.1, step0=0,
                   fixed_beta_temp=None, max_tree_depth=10, cached_lp=False, store_z=False):
        z0 = _f64(z0)
        T = self.D + self.P
        out_z = np.empty((n_iter, self.S)) if store_z else None
        tail, acc, eps = np.empty((n_iter, T)), np.empty(n_iter), np.empty(n_iter)
        nleap, depth, lp = np.empty(n_iter, dtype=np.int32), np.empty(n_iter, dtype=np.int32), np.empty(n_iter)
        self.L.mo_nuts_chain(self.h, _d(z0), int(n_iter), float(eps0), int(seed), int(chain_id),
                             int(num_adaptation_steps), float(min_temp), int(step0),
                             float("nan") if fixed_beta_temp is None else float(fixed_beta_temp), int(max_tree_depth),
                             int(bool(cached_lp)), _d(out_z) if store_z else None, _d(tail), _d(acc), _d(eps),
                             nleap.ctypes.data_as(_ip), depth.ctypes.data_as(_ip), _d(lp))
        return dict(z=out_z, tail=tail, accept=acc, step_size=eps, leapfrogs=nleap, depth=depth, lp=lp)

    def hmc_chain(self, z0, n_iter, n_leapfrog, eps0=0.1, seed=0, chain_id=0, num_adaptation_steps=0, min_temp=0.1,
                  step0=0, fixed_beta_temp=None, store_z=False):
        z0 = _f64(z0)
        T = self.D + self.P
        out_z = np.empty((n_iter, self.S)) if store_z else None
        tail, acc, eps = np.empty((n_iter, T)), np.empty(n_iter), np.empty(n_iter)
        self.L.mo_hmc_chain(self.h, _d(z0), int(n_iter), int(n_leapfrog), float(eps0), int(seed), int(chain_id),
                            int(num_adaptation_steps), float(min_temp), int(step0),
                            float("nan") if fixed_beta_temp is None else float(fixed_beta_temp),
                            _d(out_z) if store_z else None, _d(tail), _d(acc), _d(eps))
        return dict(z=out_z, tail=tail, accept=acc, step_size=eps)
