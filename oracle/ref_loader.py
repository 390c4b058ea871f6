"""Load the GENUINE reference module (/root/reference/magi_v2.py) without TensorFlow.

Test infrastructure (see oracle/__init__.py).  Only usable in the build
container: /root/reference does not exist on the GPU box, so nothing that runs
there may call this.  It is used by ``oracle/make_golden.py`` to produce the
fixtures under ``tests/golden/`` and by the CPU tests (skipped when the
reference is absent) to pin the numpy restatement in ``oracle/magi_oracle.py``.

The reference imports tensorflow / tensorflow_probability / tf_keras at module
scope (magi_v2.py:5-9) and subclasses ``tfp.mcmc.TransitionKernel``
(magi_v2.py:838).  None of those are installable here, so three stub modules
are injected into ``sys.modules``; the numpy/scipy-only methods
(_build_matrices :774, _discretize :475, _linear_interpolate :509,
cv_cubic_smoother :695) then run as the reference wrote them.
"""
import contextlib
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("MAGI_REFERENCE_DIR", "/root/reference")


class _Stub(types.ModuleType):
    """Module whose every unknown attribute is another stub (callable, no-op)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "magi_v2.py"))


def load_reference():
    """Return the imported reference module ``magi_v2`` (genuine source, stubbed TF)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_DIR}")
    if "magi_v2" in sys.modules and getattr(sys.modules["magi_v2"], "_GENUINE_REF", False):
        return sys.modules["magi_v2"]
    tf = _Stub("tensorflow")
    tf.device = lambda *_a, **_k: contextlib.nullcontext()
    tf.float64 = "float64"
    tf.int32 = "int32"
    tf.Tensor = type("Tensor", (), {})
    tf.function = lambda *a, **k: (lambda fn: fn)
    tfp = _Stub("tensorflow_probability")
    mcmc = _Stub("tensorflow_probability.mcmc")
    mcmc.TransitionKernel = type("TransitionKernel", (), {})
    tfp.mcmc = mcmc
    tfk = _Stub("tf_keras")
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "tensorflow_probability", "tf_keras")}
    sys.modules.update({"tensorflow": tf, "tensorflow_probability": tfp, "tf_keras": tfk})
    sys.path.insert(0, REFERENCE_DIR)
    try:
        # the reference sets CUDA_VISIBLE_DEVICES=-1 at import (magi_v2.py:14-16); undo it.
        cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
        mod = importlib.import_module("magi_v2")
        if cvd is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    finally:
        sys.path.remove(REFERENCE_DIR)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod._GENUINE_REF = True
    return mod


def reference_object():
    """A ``MAGI_v2`` instance created without running the TF-dependent constructor
    (magi_v2.py:53 calls tf.math.is_nan); enough to call the numpy-only methods."""
    mod = load_reference()
    return mod.MAGI_v2.__new__(mod.MAGI_v2)
