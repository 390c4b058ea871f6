"""magi_v2_b200 -- B200-native (sm_100a) implementation of the MAGI posterior-evaluation hot path of
sophiaxxiao/magi_v2: Matern covariance build, Cholesky-based factorisation, fused log-posterior +
analytic gradient, leapfrog/HMC.  Hand-written CUDA behind a C ABI (include/magi_b200.h), exposed as
``torch.ops.magi_b200.*``; Python mirrors the reference's ``MAGI_v2`` entry points.  No CPU fallback."""
import importlib

__all__ = ["MAGI_v2", "PosteriorProblem", "ops"]

_LAZY = {"ops": ("magi_v2_b200.ops", None), "PosteriorProblem": ("magi_v2_b200.ops", "PosteriorProblem"),
         "MAGI_v2": ("magi_v2_b200.magi", "MAGI_v2")}


def __getattr__(name):
    if name in _LAZY:
        mod, attr = _LAZY[name]
        m = importlib.import_module(mod)
        return m if attr is None else getattr(m, attr)
    raise AttributeError(name)
