"""ctypes binding of libmagi_b200.so (include/magi_b200.h).  No CPU fallback: if the library is
missing, loading raises; if there is no CUDA device, every compute entry point returns a CUDA
status which `check()` turns into RuntimeError."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAGI_B200_LIB", os.path.join(HERE, "libmagi_b200.so"))   # override: kernel experiments

ABI_VERSION = 3
MODEL_IDS = {"seir3": 0, "seir4": 1, "sirw": 2, "lorenz96": 3}
COV_UNIFORM_GRID = 1

c_double_p = C.c_void_p  # device pointers are passed as integers


class Problem(C.Structure):
    """magi_problem_t"""
    _fields_ = [("model_id", C.c_int), ("B", C.c_int), ("R", C.c_int), ("n", C.c_int), ("D", C.c_int),
                ("P", C.c_int), ("packed", C.c_void_p), ("mu", C.c_void_p), ("y", C.c_void_p),
                ("mask", C.c_void_p), ("N_ds", C.c_void_p), ("beta", C.c_void_p), ("LB", C.c_void_p),
                ("band", C.c_int)]


class HmcConfig(C.Structure):
    """magi_hmc_config_t"""
    _fields_ = [("n_iter", C.c_int), ("n_leapfrog", C.c_int), ("iter0", C.c_int), ("num_adapt", C.c_int),
                ("accum_from", C.c_int), ("min_temp", C.c_double), ("fixed_beta_temp", C.c_double),
                ("target_accept", C.c_double), ("seed", C.c_uint64), ("chain_id0", C.c_uint32)]


class NutsSubtree(C.Structure):
    """magi_nuts_subtree_t (include/magi_b200_nuts.h)"""
    _fields_ = [("C", C.c_int), ("nD", C.c_int), ("D", C.c_int), ("P", C.c_int)] + \
               [(k, C.c_void_p) for k in ("zc", "pc", "gc", "rho_sub", "sub_z", "sub_lp", "logw_sub", "sum_acc", "n_leaf",
                                          "building", "diverged", "ck_p", "ck_rho", "e", "H0")]


class NutsTree(C.Structure):
    """magi_nuts_tree_t (include/magi_b200_nuts.h)"""
    _fields_ = [(k, C.c_void_p) for k in ("zl", "pl", "gl", "zr", "pr", "gr", "rho", "prop_z", "prop_lp", "logw",
                                          "active", "fwd")]


_SIGNATURES = {
    "magi_b200_abi_version": (C.c_int, []),
    "magi_b200_model_dims": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "magi_b200_status_string": (C.c_char_p, [C.c_int]),
    "magi_b200_cov_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "magi_b200_factor_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "magi_b200_spd_inverse": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "magi_b200_factor_derive": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    "magi_b200_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "magi_b200_pack_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "magi_b200_sampler_workspace_bytes": (C.c_size_t, [C.POINTER(Problem)]),
    "magi_b200_logpost_grad": (C.c_int, [C.POINTER(Problem)] + [C.c_void_p] * 9 + [C.c_size_t, C.c_void_p]),
    "magi_b200_leapfrog": (C.c_int, [C.POINTER(Problem)] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.c_void_p,
                                                                                C.c_size_t, C.c_void_p]),
    "magi_b200_hmc_run": (C.c_int, [C.POINTER(Problem), C.POINTER(HmcConfig)] + [C.c_void_p] * 13 +
                          [C.c_size_t, C.c_void_p]),
    "magi_b200_logpost_grad_wide_workspace_bytes": (C.c_size_t, [C.POINTER(Problem)]),
    "magi_b200_logpost_grad_wide": (C.c_int, [C.POINTER(Problem)] + [C.c_void_p] * 9 + [C.c_size_t, C.c_void_p]),
    "magi_b200_probe_fp64": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "magi_b200_nuts_momentum": (C.c_int, [C.c_uint64, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "magi_b200_hmc_kick_drift": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_int, C.c_void_p, C.c_void_p]),
    "magi_b200_nuts_uniforms": (C.c_int, [C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "magi_b200_nuts_subtree_begin": (C.c_int, [C.POINTER(NutsSubtree), C.POINTER(NutsTree), C.c_void_p]),
    "magi_b200_nuts_merge": (C.c_int, [C.POINTER(NutsSubtree), C.POINTER(NutsTree), C.c_void_p, C.c_void_p]),
    "magi_b200_nuts_leaf_pre": (C.c_int, [C.POINTER(NutsSubtree)] + [C.c_void_p] * 5),
    "magi_b200_nuts_leaf_post": (C.c_int, [C.POINTER(NutsSubtree)] + [C.c_void_p] * 9 +
                                 [C.c_int64, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "magi_b200_nuts_leaf_post_next": (C.c_int, [C.POINTER(NutsSubtree)] + [C.c_void_p] * 9 +
                                      [C.c_int64, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python -m magi_v2_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback for the MAGI kernels.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.magi_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("libmagi_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


_user_libs = {}


def load_user_library(path: str) -> C.CDLL:
    """The library tracing.build_library compiled for a user-supplied ODE system: csrc/posterior_wide.cu with the
    generated struct, exporting the two entry points of include/magi_b200_wide.h for model id MAGI_MODEL_USER."""
    if path not in _user_libs:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: the traced f_vec has not been compiled (nvcc, sm_100a)")
        L = C.CDLL(path)
        for name in ("magi_b200_logpost_grad_wide_workspace_bytes", "magi_b200_logpost_grad_wide"):
            res, args = _SIGNATURES[name]
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _user_libs[path] = L
    return _user_libs[path]


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().magi_b200_status_string(status).decode()
        raise RuntimeError(f"{what} failed: status {status} ({msg})")
