"""User-supplied f_vec: tracing through the `tf` shim, symbolic Jacobian products, generated
device code -- everything that needs no GPU.  (The compiled operator is checked in tests/test_gpu_user_model.py.)"""
import json
import os

import numpy as np
import pytest

from magi_v2_b200 import models, tracing
from oracle.ref_loader import REFERENCE_DIR, reference_available


def seir_tf_style(t, X, thetas):
    """A user's right-hand side written with TensorFlow ops, as the reference's examples are (`tf` is resolved at call
    time: here it is not even imported -- tracing rebinds it to the shim)."""
    S = 1.0 - tf.reshape(tf.reduce_sum(X, axis=1), shape=(-1, 1))             # noqa: F821
    dE = thetas[0] * S * X[:, 1:2] - thetas[2] * X[:, 0:1]
    dI = thetas[2] * X[:, 0:1] - thetas[1] * X[:, 1:2]
    return tf.concat([dE, dI, thetas[1] * X[:, 1:2]], axis=1)                 # noqa: F821


def saturating_sir(t, X, th):
    """Not affine in theta, uses an elementary function; numpy ops."""
    S, I = X[:, 0:1], X[:, 1:2]
    inc = th[0] * S * I / (1.0 + th[2] * I)
    return np.concatenate([-inc, inc - th[1] * np.exp(-0.1 * I) * I], axis=1)


def test_trace_tf_style_callable_matches_registry_system():
    ts = tracing.trace(seir_tf_style, 3, 3)
    assert ts.affine_in_theta
    rng = np.random.default_rng(1)
    X, th = rng.uniform(0, 1, (9, 3)), rng.uniform(0.1, 2, 3)
    ref = models.REGISTRY["seir3"]
    assert np.allclose(ts.numpy_f()(None, X, th), ref.f_vec(None, X, th), rtol=1e-14, atol=1e-15)
    assert np.allclose(ts.numpy_dtheta()(None, X, th), ref.dtheta(None, X, th), rtol=1e-14, atol=1e-15)
    # without jit=True a callable that reproduces a compiled-in system resolves to it (fused kernels)
    assert models.resolve(seir_tf_style, 3, 3, jit=False) is ref


def test_generated_vjp_matches_finite_differences():
    import sympy
    ts = tracing.trace(saturating_sir, 2, 3)
    assert not ts.affine_in_theta
    code = tracing.emit_cuda(ts)
    assert "struct UserModel" in code and "D = 2, P = 3" in code and "exp(" in code
    # the emitted expressions, evaluated by sympy, against central differences of the callable
    gs = [sympy.Symbol(f"g{d}", real=True) for d in range(2)]
    vx = [sum(gs[dp] * sympy.diff(ts.f[dp], ts.x[d]) for dp in range(2)) for d in range(2)]
    vth = [sum(gs[dp] * sympy.diff(ts.f[dp], ts.th[k]) for dp in range(2)) for k in range(3)]
    lam = sympy.lambdify([ts.x, ts.th, gs], vx + vth, modules="numpy")
    rng = np.random.default_rng(2)
    x, th, g = rng.uniform(0.1, 0.9, 2), rng.uniform(0.2, 1.5, 3), rng.standard_normal(2)
    got = np.array(lam(list(x), list(th), list(g)), dtype=np.float64)
    f = lambda xx, tt: saturating_sir(None, xx[None], tt)[0]
    h = 1e-6
    fd = []
    for d in range(2):
        e = np.zeros(2); e[d] = h
        fd.append(g @ (f(x + e, th) - f(x - e, th)) / (2 * h))
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd.append(g @ (f(x, th + e) - f(x, th - e)) / (2 * h))
    assert np.allclose(got, fd, rtol=1e-7, atol=1e-9)


def test_tracer_rejects_what_the_kernels_cannot_run():
    with pytest.raises(NotImplementedError):                                    # depends on t
        tracing.trace(lambda t, X, th: X * th[0] + t, 2, 1)
    with pytest.raises(ValueError):                                             # couples grid points
        tracing.trace(lambda t, X, th: np.cumsum(X, axis=0) * th[0], 2, 1)
    with pytest.raises(ValueError):                                             # wrong output shape
        tracing.trace(lambda t, X, th: X[:, 0:1] * th[0], 2, 1)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_the_references_own_callables_trace_unmodified():
    """vignette.ipynb:68-79 and test_magi_script.py:19-45, source taken from the reference as it is (with its tf.* calls)."""
    nb = json.load(open(os.path.join(REFERENCE_DIR, "vignette.ipynb")))
    ns = {}
    exec("".join(nb["cells"][3]["source"]), ns)
    ts = tracing.trace(ns["f_vec"], 3, 3)
    rng = np.random.default_rng(3)
    X, th = rng.uniform(0, 1, (9, 3)), rng.uniform(0.1, 2, 3)
    assert np.allclose(ts.numpy_f()(None, X, th), models.REGISTRY["seir3"].f_vec(None, X, th), rtol=1e-14, atol=1e-15)
    src = open(os.path.join(REFERENCE_DIR, "test_magi_script.py")).read()
    a = src.index("def f_vec")
    b = src.index("\n\n", src.index("return", a))
    ns = {}
    exec(src[a:b], ns)
    ts = tracing.trace(ns["f_vec"], 4, 5)
    X, th = rng.uniform(0, 1, (9, 4)), rng.uniform(0.1, 2, 5)
    assert np.allclose(ts.numpy_f()(None, X, th), models.REGISTRY["sirw"].f_vec(None, X, th), rtol=1e-14, atol=1e-15)
