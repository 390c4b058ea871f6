"""Host-side registry of the ODE systems compiled into libmagi_b200.so (csrc/ode_models.cuh).

The reference takes the ODE as a Python/TF callable ``f_vec(t, X[n,D], thetas[P]) -> [n,D]``
(magi_v2.py:32-33, :73).  The fused CUDA kernels need f and its Jacobian products as device code, so
the drop-in accepts (a) a registry name, (b) an ``OdeModel``, or (c) a callable that is *identified*
against the registry by probing it with numpy inputs (``resolve``).  The numpy forms below are used
only for host-side initialisation (theta init, data synthesis), never on the sampling path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Union

import numpy as np


def _cat(cols, like):
    """Column concatenation for numpy arrays or torch tensors (the latter only for the autograd-based
    initialisation of unobserved components, init_fit.py)."""
    if isinstance(like, np.ndarray):
        return np.concatenate(cols, axis=1)
    import torch
    return torch.cat(cols, dim=1)


def _roll(X, k):
    if isinstance(X, np.ndarray):
        return np.roll(X, k, axis=1)
    import torch
    return torch.roll(X, k, dims=1)


def _seir3(t, X, th):
    S = 1.0 - X.sum(1)[:, None]
    E, I = X[:, 0:1], X[:, 1:2]
    return _cat([th[0] * S * I - th[2] * E, th[2] * E - th[1] * I, th[1] * I], X)


def _seir4(t, X, th):
    S, E, I = X[:, 0:1], X[:, 1:2], X[:, 2:3]
    return _cat([-th[0] * S * I, th[0] * S * I - th[2] * E, th[2] * E - th[1] * I, th[1] * I], X)


def _sirw(t, X, th):
    S, I, R, W = X[:, 0:1], X[:, 1:2], X[:, 2:3], X[:, 3:4]
    beta, phi, xi, chi, kappa = th[0], th[1], th[2], th[3], th[4]
    return _cat([-beta * S * I + kappa * W, beta * S * I - phi * I, phi * I - xi * R + chi * I * W,
                 xi * R - chi * I * W - kappa * W], X)


def _lorenz96(t, X, th):
    return (_roll(X, -1) - _roll(X, 2)) * _roll(X, 1) - X + th[0]


def _dtheta_fd(f):
    """d f / d theta [n, D, P] -- every registry right-hand side is affine in theta, so a unit
    difference is exact."""
    def g(t, X, th):
        th = np.asarray(th, dtype=np.float64)
        f0 = f(t, X, np.zeros_like(th))
        return np.stack([f(t, X, np.eye(len(th))[k]) - f0 for k in range(len(th))], axis=2)
    return g


@dataclass(frozen=True)
class OdeModel:
    name: str
    model_id: int          # magi_model_t in include/magi_b200.h
    D: int
    P: int
    f_vec: Callable        # numpy (t, X[n,D], th[P]) -> [n,D]
    dtheta: Callable       # numpy (t, X, th) -> [n,D,P]
    components: tuple
    parameters: tuple
    lib_path: str = None   # a user system (model_id MAGI_MODEL_USER): the library tracing.build_library compiled for it
    affine_in_theta: bool = True


REGISTRY: Dict[str, OdeModel] = {
    "seir3": OdeModel("seir3", 0, 3, 3, _seir3, _dtheta_fd(_seir3), ("E", "I", "R"), ("beta", "gamma", "sigma")),
    "seir4": OdeModel("seir4", 1, 4, 3, _seir4, _dtheta_fd(_seir4), ("S", "E", "I", "R"), ("beta", "gamma", "sigma")),
    "sirw": OdeModel("sirw", 2, 4, 5, _sirw, _dtheta_fd(_sirw), ("S", "I", "R", "W"),
                     ("beta", "phi", "xi", "chi", "kappa")),
    "lorenz96": OdeModel("lorenz96", 3, 10, 1, _lorenz96, _dtheta_fd(_lorenz96),
                         tuple(f"x{i}" for i in range(10)), ("F",)),
}


def user_model(f_vec: Callable, D: int, P: int) -> OdeModel:
    """Trace a user right-hand side, generate its device code and compile its library (tracing.py; row f4)."""
    from . import tracing
    ts = tracing.trace(f_vec, D, P)
    so = tracing.build_library(ts)
    return OdeModel(f"user_{ts.source_hash}", tracing.MODEL_USER, D, P, ts.numpy_f(), ts.numpy_dtheta(),
                    tuple(f"x{i}" for i in range(D)), tuple(f"theta{k}" for k in range(P)), lib_path=so,
                    affine_in_theta=ts.affine_in_theta)


def resolve(f_vec: Union[str, OdeModel, Callable], D: int, P: int, jit: Union[bool, None] = None) -> OdeModel:
    """Map the constructor's ``f_vec`` argument to device code: a compiled-in system (name, OdeModel, or a callable
    that reproduces one when probed), else -- or always, with jit=True -- a traced and run-time-compiled user system.
    A callable written with ``tf.*`` ops is evaluated through the shim of tracing.py (TensorFlow itself is not needed)."""
    if isinstance(f_vec, OdeModel):
        return f_vec
    if isinstance(f_vec, str):
        if f_vec not in REGISTRY:
            raise KeyError(f"unknown ODE model {f_vec!r}; compiled-in models: {sorted(REGISTRY)}")
        return REGISTRY[f_vec]
    if callable(f_vec):
        if jit:
            return user_model(f_vec, D, P)
        from . import tracing
        rng = np.random.default_rng(0)
        X = rng.uniform(0.05, 0.5, (7, D))
        th = rng.uniform(0.1, 2.0, P)
        t = np.linspace(0, 1, 7).reshape(-1, 1)
        try:
            out = np.asarray(tracing._rebind(f_vec)(t, X, th), dtype=np.float64)
        except Exception:  # noqa: BLE001  (ops outside the shim: let the tracer give the precise error)
            out = None
        if out is not None:
            for m in REGISTRY.values():
                if (m.D, m.P) == (D, P) and out.shape == (7, D) and np.allclose(out, m.f_vec(t, X, th), rtol=1e-12,
                                                                                 atol=1e-14):
                    return m
        if jit is False:
            raise ValueError("f_vec does not match any ODE system compiled into libmagi_b200.so "
                             f"(D={D}, P={P}); available: {sorted(REGISTRY)}")
        return user_model(f_vec, D, P)
    raise TypeError("f_vec must be a model name, an OdeModel, or a callable")
