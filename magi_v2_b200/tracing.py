"""User-supplied ODE right-hand sides.

The reference takes the ODE as a Python callable written with TensorFlow ops,
``f_vec(t, X[n, D], thetas[P]) -> [n, D]`` (magi_v2.py:32-33, :73; called at :155, :206, :335; examples:
vignette.ipynb:68-79, test_magi_script.py:19-45), and obtains its Jacobians by TF autodiff.  The CUDA kernels need
``f`` and its vector-Jacobian products as device code, so a callable that is not one of the compiled-in systems is

1. **traced**: called once on a single grid point with ``X`` and ``thetas`` made of sympy symbols (numpy object
   arrays), with the name ``tf`` in the callable's globals rebound to the small shim below (``reshape``, ``concat``,
   ``reduce_sum``, slicing, arithmetic, elementary functions ... -- every op is pointwise in time), which yields the D
   expressions f_d(x, theta);
2. **differentiated** symbolically: the products  vx[d] = sum_d' g[d'] df_d'/dx_d  and
   vth[k] = sum_d' g[d'] df_d'/dtheta_k  that reverse-mode autodiff of ``f_vec`` contributes at :335;
3. **emitted** as ``struct UserModel`` in the form of csrc/ode_models.cuh and **compiled** with nvcc (sm_100a) into
   a library of its own from csrc/posterior_wide.cu -- the log-posterior + gradient operator for that system
   (``magi_b200_logpost_grad_wide``, include/magi_b200_wide.h; model id MAGI_MODEL_USER).  The NUTS leaf kernels and
   the host-driven HMC are model-agnostic and run on top of it.

The traced expressions are checked against the callable itself on random inputs before anything is compiled.
There is no CPU fallback: without nvcc / a CUDA device the build or the first launch raises."""
from __future__ import annotations

import hashlib
import os
import subprocess
import types
from dataclasses import dataclass
from typing import Callable, List

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
JIT_DIR = os.path.join(HERE, "jit")
MODEL_USER = 100                      # MAGI_MODEL_USER, include/magi_b200.h


# --------------------------------------------------------------------------------------------------------------------
# the `tf` stand-in: numpy semantics, works on float arrays and on object arrays of sympy expressions alike
# --------------------------------------------------------------------------------------------------------------------
class _E:
    """One symbolic scalar: a sympy expression with Python arithmetic and the method names numpy's ufuncs look up on
    the elements of object arrays (``np.exp(a)`` calls ``a[i].exp()``), so that numpy / the `tf` shim below act on
    arrays of them exactly as on float arrays."""
    __slots__ = ("e",)

    def __init__(self, e):
        self.e = e

    @staticmethod
    def _w(o):
        return o.e if isinstance(o, _E) else o

    def __add__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(self.e + _E._w(o))
    def __radd__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(_E._w(o) + self.e)
    def __sub__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(self.e - _E._w(o))
    def __rsub__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(_E._w(o) - self.e)
    def __mul__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(self.e * _E._w(o))
    def __rmul__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(_E._w(o) * self.e)
    def __truediv__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(self.e / _E._w(o))
    def __rtruediv__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(_E._w(o) / self.e)
    def __pow__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(self.e ** _E._w(o))
    def __rpow__(self, o): return NotImplemented if isinstance(o, np.ndarray) else _E(_E._w(o) ** self.e)
    def __neg__(self): return _E(-self.e)
    def __pos__(self): return self
    def square(self): return _E(self.e * self.e)
    def reciprocal(self): return _E(1 / self.e)


def _unary(name):
    def method(self):
        import sympy
        return _E(getattr(sympy, name)(self.e))
    return method


for _nm in ("exp", "log", "sqrt", "sin", "cos", "tan", "tanh", "sinh", "cosh"):
    setattr(_E, _nm, _unary(_nm))


def _elementwise(name):
    def fn(x, *a, **k):
        return getattr(np, name)(np.asarray(x))
    return fn


def _make_tf_shim():
    ns = types.SimpleNamespace()
    ns.float64, ns.float32, ns.int32, ns.int64 = np.float64, np.float32, np.int32, np.int64
    ns.newaxis = None
    ns.reshape = lambda tensor, shape=None, name=None: np.reshape(tensor, shape)
    ns.concat = lambda values, axis=0, name=None: np.concatenate([np.asarray(v) for v in values], axis=axis)
    ns.stack = lambda values, axis=0, name=None: np.stack([np.asarray(v) for v in values], axis=axis)
    ns.transpose = lambda a, perm=None, **k: np.transpose(a, perm)
    ns.expand_dims = lambda a, axis, **k: np.expand_dims(a, axis)
    ns.squeeze = lambda a, axis=None, **k: np.squeeze(a, axis)
    ns.reduce_sum = lambda a, axis=None, keepdims=False, **k: np.sum(a, axis=axis, keepdims=keepdims)
    ns.reduce_mean = lambda a, axis=None, keepdims=False, **k: np.mean(a, axis=axis, keepdims=keepdims)
    ns.reduce_prod = lambda a, axis=None, keepdims=False, **k: np.prod(a, axis=axis, keepdims=keepdims)
    ns.cast = lambda x, dtype=None, **k: x
    ns.convert_to_tensor = lambda x, dtype=None, **k: np.asarray(x)
    ns.constant = lambda x, dtype=None, **k: np.asarray(x)
    ns.identity = lambda x, **k: x
    ns.stop_gradient = lambda x, **k: x
    ns.ones_like = lambda x, **k: np.ones(np.shape(x))
    ns.zeros_like = lambda x, **k: np.zeros(np.shape(x))
    ns.ones = lambda shape, dtype=None, **k: np.ones(shape)
    ns.zeros = lambda shape, dtype=None, **k: np.zeros(shape)
    ns.shape = lambda x, **k: np.shape(x)
    ns.gather = lambda params, indices, axis=0, **k: np.take(params, indices, axis=axis)
    ns.roll = lambda x, shift, axis, **k: np.roll(x, shift, axis=axis)
    ns.add, ns.subtract, ns.multiply, ns.divide = (lambda a, b, **k: a + b), (lambda a, b, **k: a - b), \
        (lambda a, b, **k: a * b), (lambda a, b, **k: a / b)
    ns.pow = lambda a, b, **k: a ** b
    ns.square = lambda a, **k: a * a
    ns.negative = lambda a, **k: -a
    for nm in ("exp", "log", "sqrt", "sin", "cos", "tan", "tanh", "sinh", "cosh"):
        setattr(ns, nm, _elementwise(nm))
    ns.sigmoid = lambda x, **k: 1.0 / (1.0 + ns.exp(-np.asarray(x)))
    ns.math = ns
    ns.linalg = types.SimpleNamespace(matmul=lambda a, b, **k: np.matmul(a, b))
    ns.device = lambda *a, **k: __import__("contextlib").nullcontext()
    ns.function = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda fn: fn))
    return ns


tf = _make_tf_shim()


def _rebind(f_vec: Callable) -> Callable:
    """The same code object with the global names `tf` / `tensorflow` bound to the shim (the callable is not modified)."""
    if not isinstance(f_vec, types.FunctionType):
        return f_vec
    g = dict(f_vec.__globals__)
    g["tf"] = tf
    g["tensorflow"] = tf
    out = types.FunctionType(f_vec.__code__, g, f_vec.__name__, f_vec.__defaults__, f_vec.__closure__)
    out.__kwdefaults__ = f_vec.__kwdefaults__
    return out


# --------------------------------------------------------------------------------------------------------------------
# tracing
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class TracedSystem:
    D: int
    P: int
    f: list                    # D sympy expressions in x0..x{D-1}, th0..th{P-1}
    x: list
    th: list
    affine_in_theta: bool
    source_hash: str

    def numpy_f(self):
        import sympy
        fn = sympy.lambdify([self.x, self.th], self.f, modules="numpy")
        D = self.D

        def f_vec(t, X, thetas):
            cols = fn([X[:, d] for d in range(D)], list(thetas))
            like = X[:, 0]
            return _stack_cols([c + 0 * like for c in cols], X)          # constants broadcast to the grid

        return f_vec

    def numpy_dtheta(self):
        import sympy
        J = [[sympy.diff(fd, tk) for tk in self.th] for fd in self.f]
        fn = sympy.lambdify([self.x, self.th], J, modules="numpy")
        D, P = self.D, self.P

        def dtheta(t, X, thetas):
            vals = fn([X[:, d] for d in range(D)], list(np.asarray(thetas, dtype=np.float64)))
            out = np.empty((X.shape[0], D, P))
            for d in range(D):
                for k in range(P):
                    out[:, d, k] = vals[d][k]
            return out

        return dtheta


def _stack_cols(cols, like):
    if isinstance(like, np.ndarray):
        return np.stack(cols, axis=1)
    import torch
    return torch.stack(cols, dim=1)


def trace(f_vec: Callable, D: int, P: int) -> TracedSystem:
    """Symbolic form of a user right-hand side; raises if it is not pointwise in time, depends on t, or uses an op the
    shim does not know."""
    import sympy
    xs = [sympy.Symbol(f"x{d}", real=True) for d in range(D)]
    ths = [sympy.Symbol(f"th{k}", real=True) for k in range(P)]
    tt = sympy.Symbol("t_", real=True)
    fn = _rebind(f_vec)
    X = np.empty((1, D), dtype=object)
    th = np.empty((P,), dtype=object)
    tg = np.empty((1, 1), dtype=object)
    for d in range(D):
        X[0, d] = _E(xs[d])
    for k in range(P):
        th[k] = _E(ths[k])
    tg[0, 0] = _E(tt)
    try:
        out = np.asarray(fn(tg, X, th), dtype=object)
    except Exception as e:  # noqa: BLE001
        raise TypeError("f_vec could not be traced: it must be written with tf.* / numpy ops that act pointwise in "
                        f"time (reshape, concat, reduce_sum over components, slicing, arithmetic); {type(e).__name__}: {e}") from e
    if out.shape != (1, D):
        raise ValueError(f"f_vec must return [n, D] = [n, {D}]; traced shape {out.shape}")
    f = [sympy.sympify(_E._w(out[0, d])) for d in range(D)]
    if any(tt in fd.free_symbols for fd in f):
        raise NotImplementedError("non-autonomous systems (f depending on t) are not supported by the CUDA path")
    extra = set().union(*[fd.free_symbols for fd in f]) - set(xs) - set(ths)
    if extra:
        raise ValueError(f"f_vec produced unknown symbols {extra}")
    # the traced expressions reproduce the callable on a whole grid (i.e. it really is pointwise in time)
    rng = np.random.default_rng(0)
    Xn, thn, tn = rng.uniform(0.05, 0.9, (7, D)), rng.uniform(0.1, 2.0, P), np.linspace(0, 1, 7).reshape(-1, 1)
    want = np.asarray(fn(tn, Xn, thn), dtype=np.float64)
    lam = sympy.lambdify([xs, ths], f, modules="numpy")
    got = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), (7,)) for c in lam([Xn[:, d] for d in range(D)], list(thn))], axis=1)
    if want.shape != (7, D) or not np.allclose(got, want, rtol=1e-12, atol=1e-14):
        raise ValueError("f_vec is not pointwise in time (row i of the output must depend on row i of X only)")
    affine = all(sympy.simplify(sympy.diff(fd, a, b)) == 0 for fd in f for a in ths for b in ths)
    h = hashlib.sha256(("|".join(sympy.srepr(fd) for fd in f) + f"|{D}|{P}").encode()).hexdigest()[:16]
    return TracedSystem(D, P, f, xs, ths, affine, h)


# --------------------------------------------------------------------------------------------------------------------
# code generation
# --------------------------------------------------------------------------------------------------------------------
def emit_cuda(ts: TracedSystem) -> str:
    """`struct UserModel` with the interface of the structs in csrc/ode_models.cuh."""
    import sympy
    from sympy.printing.c import C99CodePrinter

    class Printer(C99CodePrinter):
        def _print_Pow(self, expr):
            b, e = expr.as_base_exp()
            if e.is_Integer and 2 <= int(e) <= 4:
                return "(" + "*".join([self._print(b) if b.is_Atom else "(" + self._print(b) + ")"] * int(e)) + ")"
            if e.is_Integer and -4 <= int(e) <= -1:
                den = "*".join([self._print(b) if b.is_Atom else "(" + self._print(b) + ")"] * (-int(e)))
                return f"(1.0/({den}))"
            return super()._print_Pow(expr)

    pr = Printer()
    gs = [sympy.Symbol(f"g{d}", real=True) for d in range(ts.D)]
    vx = [sum(gs[dp] * sympy.diff(ts.f[dp], ts.x[d]) for dp in range(ts.D)) for d in range(ts.D)]
    vth = [sum(gs[dp] * sympy.diff(ts.f[dp], ts.th[k]) for dp in range(ts.D)) for k in range(ts.P)]
    names = {**{s: f"x[{i}]" for i, s in enumerate(ts.x)}, **{s: f"th[{i}]" for i, s in enumerate(ts.th)},
             **{s: f"g[{i}]" for i, s in enumerate(gs)}}

    def body(exprs: List, outs: List[str]) -> str:
        repl, red = sympy.cse(exprs, symbols=sympy.numbered_symbols("c_"), optimizations="basic")
        lines = []
        for sym, e in repl:
            lines.append(f"    const double {sym} = {pr.doprint(e)};")
        for o, e in zip(outs, red):
            lines.append(f"    {o} = {pr.doprint(e)};")
        txt = "\n".join(lines)
        for s, nm in names.items():
            txt = _replace_symbol(txt, str(s), nm)
        return txt

    f_body = body(list(ts.f), [f"out[{d}]" for d in range(ts.D)])
    v_body = body(vx + vth, [f"vx[{d}]" for d in range(ts.D)] + [f"vth[{k}]" for k in range(ts.P)])
    return f"""// generated by magi_v2_b200/tracing.py from a user-supplied f_vec (hash {ts.source_hash}); do not edit
#pragma once
struct UserModel {{
  static constexpr int D = {ts.D}, P = {ts.P}, ID = MAGI_MODEL_USER;
  __device__ static __forceinline__ void f(const double* x, const double* th, double* out) {{
{f_body}
  }}
  __device__ static __forceinline__ void vjp(const double* x, const double* th, const double* g, double* vx,
                                             double* vth) {{
{v_body}
  }}
}};
"""


def _replace_symbol(txt: str, sym: str, repl: str) -> str:
    import re
    return re.sub(rf"(?<![A-Za-z0-9_]){re.escape(sym)}(?![A-Za-z0-9_])", repl, txt)


# --------------------------------------------------------------------------------------------------------------------
# run-time compilation
# --------------------------------------------------------------------------------------------------------------------
def build_library(ts: TracedSystem, verbose: bool = False) -> str:
    """nvcc (sm_100a) csrc/posterior_wide.cu with the generated struct -> magi_v2_b200/jit/<hash>/libmagi_user.so"""
    d = os.path.join(JIT_DIR, ts.source_hash)
    so = os.path.join(d, "libmagi_user.so")
    hdr = os.path.join(d, "user_model.cuh")
    code = emit_cuda(ts)
    deps = [os.path.join(HERE, "csrc", f) for f in ("posterior_wide.cu", "common.cuh", "ode_models.cuh")] + \
           [os.path.join(os.path.dirname(HERE), "include", f) for f in ("magi_b200.h", "magi_b200_wide.h")]
    newest = max(os.path.getmtime(f) for f in deps if os.path.exists(f))
    if os.path.exists(so) and os.path.exists(hdr) and open(hdr).read() == code and os.path.getmtime(so) >= newest:
        return so      # (a library older than the sources it was compiled from -- e.g. a changed packed layout -- is rebuilt)
    os.makedirs(d, exist_ok=True)
    with open(hdr, "w") as f:
        f.write(code)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-Xcompiler", "-fvisibility=hidden", f'-DMAGI_USER_MODEL_HEADER="{hdr}"', "-shared", "-o", so,
           os.path.join(HERE, "csrc", "posterior_wide.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for the traced f_vec:\n{r.stdout}\n{r.stderr}\n--- generated code ---\n{code}")
    if verbose:
        print(f"[magi_v2_b200.tracing] built {so}")
    return so
