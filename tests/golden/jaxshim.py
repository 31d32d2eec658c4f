"""Run the reference's OWN Python sources (read in place from /root/reference) without JAX.

TEST INFRASTRUCTURE, used only by tests/golden/make_reference_golden.py (fixture generation in the build container) and
by the optional live check in tests/test_oracle_vs_reference.py.  Nothing under sparsernns_b200/ imports it.

The reference is pure Python over ``jax.numpy`` / ``jax.lax`` / ``flax``; none of them is installable here (DESIGN.md
section 2).  Its arithmetic on the hot path, however, only touches a small part of ``jax.numpy`` plus ``jax.vmap``,
``jax.lax.scan``, ``jax.lax.associative_scan``, ``jax.lax.rsqrt``, ``jax.nn.relu`` and ``jax.tree_util``.  ``install()``
registers stand-ins for those, backed by NumPy, with JAX's default 32-bit semantics:

* every result is demoted float64 -> float32, int64 -> int32, complex128 -> complex64 (JAX without x64 has no 64-bit
  types; ``x.astype(np.int64)`` is a no-op on int32 data, fxparray.py:613-616);
* an operation that NumPy would carry out in float64 / complex128 only because an int32 array met a Python float /
  complex (``x * 2.0**-e``, ``re + 1j * im``) is redone with the integers converted to float32 first, as JAX's weak
  typing does (this is what makes the float32 round trip of fxp_relu, fxpmodel.py:30-45, lossy above 2**24);
* int32 arithmetic wraps, ``>>`` is arithmetic, ``round`` is half-to-even - NumPy and XLA agree on these.

``flax``, ``aqt`` and ``jaxtyping`` are replaced by inert stubs: the reference's modules import them at module level
(class definitions, decorators, type annotations), but the code executed here - fxparray.py, fxpmodel.py, fxputils.py
and the pure functions of ssm.py / quantization.py - never calls into them.

What this does and does not pin: it executes the reference's control flow, shifts, rounding, saturation, exponent
bookkeeping and operator order exactly as written, so it pins the oracle as a RESTATEMENT of that source.  It is not
XLA: transcendental functions (log2, exp, sqrt) are NumPy's, so last-ulp differences of XLA's float32 ``log2`` at
``compute_best`` boundaries (fxparray.py:424, :606) remain unpinned.
"""

from __future__ import annotations

import functools
import sys
import types

import numpy as _np

REFERENCE_ROOT = "/root/reference"

_DEMOTE = {_np.dtype("float64"): _np.float32, _np.dtype("int64"): _np.int32, _np.dtype("uint64"): _np.uint32,
           _np.dtype("complex128"): _np.complex64}


class JArr(_np.ndarray):
    """ndarray with JAX's 32-bit result types; scalars stay 0-d arrays (as jax.Array does)."""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        args = [_raw(x) for x in inputs]
        if out is not None:
            kw["out"] = tuple(_raw(o) for o in out)
        res = getattr(ufunc, method)(*args, **kw)
        if _is64(res) and method in ("__call__", "reduce", "accumulate", "outer"):
            wide = _np.complexfloating if _kind(res) == "c" else _np.floating
            if _kind(res) in "fc" and not any(isinstance(a, _np.ndarray) and a.dtype in (_np.float64, _np.complex128)
                                              for a in args):
                # int32 array (op) Python float/complex: JAX converts the integers to float32 first
                args = [a.astype(_np.float32) if isinstance(a, (_np.ndarray, _np.generic)) and a.dtype.kind in "iub"
                        else a for a in args]
                res = getattr(ufunc, method)(*args, **kw)
            del wide
        return _demote(res)

    def __getitem__(self, idx):
        return _demote(super().__getitem__(idx))

    def astype(self, dtype, *a, **k):
        return _demote(_np.asarray(self).astype(_np.dtype(dtype), *a, **k))

    # jax.Array API used by the reference
    @property
    def at(self):
        return _At(self)

    def item(self, *a):
        return _np.asarray(self).item(*a)

    def __hash__(self):
        return id(self)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Upd:
            def set(self, v):
                out = _np.array(arr)
                out[idx] = _raw(v)
                return _demote(out)

            def add(self, v):
                out = _np.array(arr)
                out[idx] += _raw(v)
                return _demote(out)

        return _Upd()


def _kind(x):
    return _np.asarray(x).dtype.kind


def _is64(x):
    if isinstance(x, tuple):
        return any(_is64(e) for e in x)
    return isinstance(x, (_np.ndarray, _np.generic)) and x.dtype in _DEMOTE


def _raw(x):
    if isinstance(x, JArr):
        return x.view(_np.ndarray)
    if isinstance(x, (list, tuple)):
        return type(x)(_raw(e) for e in x)
    return x


def _demote(x):
    if isinstance(x, (tuple, list)):
        return type(x)(_demote(e) for e in x)
    if isinstance(x, _np.generic):
        x = _np.asarray(x)
    if isinstance(x, _np.ndarray):
        t = _DEMOTE.get(x.dtype)
        if t is not None:
            x = x.astype(t)
        return x.view(JArr)
    return x


def _wrap(fn):
    @functools.wraps(fn)
    def call(*a, **k):
        if "dtype" in k and k["dtype"] is not None:
            k["dtype"] = _DEMOTE.get(_np.dtype(k["dtype"]), k["dtype"])
        return _demote(fn(*[_raw(x) for x in a], **{n: _raw(v) for n, v in k.items()}))

    return call


class _NumpyShim(types.ModuleType):
    """``jax.numpy``: NumPy functions with 32-bit results."""

    int64 = _np.int32          # jnp.int64 silently is int32 without jax_enable_x64
    float64 = _np.float32
    complex128 = _np.complex64
    ndarray = _np.ndarray

    def __getattr__(self, name):
        obj = getattr(_np, name)
        if isinstance(obj, type) or not callable(obj):
            return obj
        w = _wrap(obj)
        setattr(self, name, w)
        return w


def _float_fn(fn):
    """Transcendentals: JAX evaluates them in float32 (integers are converted to float32 first)."""

    def call(x, *a, **k):
        x = _np.asarray(_raw(x))
        if x.dtype.kind in "iub" or x.dtype == _np.float64:
            x = x.astype(_np.float32)
        elif x.dtype == _np.complex128:
            x = x.astype(_np.complex64)
        return _demote(fn(x, *a, **k))

    return call


def _make_jnp():
    m = _NumpyShim("jax.numpy")
    for name in ("log2", "log", "exp", "sqrt", "tanh", "sin", "cos", "log10", "exp2"):
        setattr(m, name, _float_fn(getattr(_np, name)))
    m.array = _wrap(_np.array)
    m.asarray = _wrap(_np.asarray)
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.eigh = _wrap(_np.linalg.eigh)
    m.linalg = linalg
    return m, linalg


# ---------------------------------------------------------------------------------------------------------------------
# jax.*: tree utilities, vmap, lax.scan, lax.associative_scan (sequential fold - results within float tolerance of
# XLA's tree order, identical for exact integer operators), nn.relu (lexicographic on complex, like jnp.maximum)
def _tree_map(f, t, *rest):
    if isinstance(t, dict):
        return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return type(t)(_tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t))
    return f(t, *rest)


class _Key:
    def __init__(self, key):
        self.key = key

    def __str__(self):
        return f"['{self.key}']"


def _tree_map_with_path(f, t, path=()):
    if isinstance(t, dict):
        return {k: _tree_map_with_path(f, v, path + (_Key(k),)) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return type(t)(_tree_map_with_path(f, v, path + (_Key(i),)) for i, v in enumerate(t))
    return f(path, t)


def _tree_leaves(t):
    out = []
    _tree_map(out.append, t)
    return out


def _vmap(fn, in_axes=0, out_axes=0, **_unused):
    def call(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(_tree_leaves(a)[0])[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else _tree_map(lambda v: _demote(_np.take(_raw(v), i, axis=ax)), a)
                  for a, ax in zip(args, axes)]
            outs.append(fn(*sl))
        return _tree_map(lambda *vs: _demote(_np.stack([_raw(v) for v in vs], axis=out_axes)), outs[0], *outs[1:])

    return call


def _scan(step, init, xs, length=None, reverse=False, unroll=1):
    n = length if xs is None else _np.shape(_tree_leaves(xs)[0])[0]
    carry = init
    ys = []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        x_i = None if xs is None else _tree_map(lambda v: _demote(_raw(v)[i]), xs)
        carry, y = step(carry, x_i)
        ys.append(y)
    if reverse:
        ys.reverse()
    stacked = None if ys[0] is None else _tree_map(lambda *vs: _demote(_np.stack([_raw(v) for v in vs])), ys[0], *ys[1:])
    return carry, stacked


def _associative_scan(fn, elems, reverse=False, axis=0):
    """Sequential left fold with the caller's (vmapped) binary operator: out[k] = fn(out[k-1], e[k])."""
    assert axis == 0
    n = _np.shape(_tree_leaves(elems)[0])[0]
    order = range(n - 1, -1, -1) if reverse else range(n)
    outs = [None] * n
    prev = None
    for i in order:
        e_i = _tree_map(lambda v: _demote(_raw(v)[i:i + 1]), elems)
        prev = e_i if prev is None else fn(prev, e_i)
        outs[i] = prev
    return _tree_map(lambda *vs: _demote(_np.concatenate([_raw(v) for v in vs])), outs[0], *outs[1:])


class _Stub(types.ModuleType):
    """Inert stand-in for flax / aqt / jaxtyping: any attribute exists; CamelCase names are classes, the rest are
    pass-through decorators / factories."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            obj = _StubMeta(name, (), {})
        else:
            def obj(*a, **k):
                if len(a) >= 1 and callable(a[0]) and not k:
                    return a[0]            # used as a decorator / class transform (nn.compact, nn.vmap(cls, ...))
                return _Stub(name)
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        return a[0] if a and callable(a[0]) else self

    def __getitem__(self, item):
        return self


class _StubMeta(type):
    def __getitem__(cls, item):
        return cls

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(name)


def install(reference_root: str = REFERENCE_ROOT):
    """Register the stand-ins in sys.modules and put the reference on sys.path.  Idempotent."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_s5b200_shim", False):
        return sys.modules["jax"]
    if "jax" in sys.modules:
        raise RuntimeError("a real jax is already imported; the shim is only for containers without JAX")
    jnp, linalg = _make_jnp()
    jax = types.ModuleType("jax")
    jax._s5b200_shim = True
    jax.numpy = jnp
    jax.Array = _np.ndarray
    jax.vmap = _vmap
    jax.jit = lambda f=None, **k: f if f is not None else (lambda g: g)
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.tree_map = _tree_map
    tree_util.tree_map_with_path = _tree_map_with_path
    tree_util.tree_leaves = _tree_leaves
    jax.tree_util = tree_util
    jax.tree_map = _tree_map
    lax = types.ModuleType("jax.lax")
    lax.scan = _scan
    lax.associative_scan = _associative_scan
    lax.rsqrt = _float_fn(lambda x: (1.0 / _np.sqrt(x)).astype(x.dtype))
    lax.stop_gradient = lambda x: x
    jax.lax = lax
    nn = _Stub("jax.nn")
    nn.relu = _wrap(lambda x: _np.maximum(x, 0))
    nn.sigmoid = _float_fn(lambda x: (1 / (1 + _np.exp(-x))).astype(x.dtype))
    nn.initializers = _Stub("jax.nn.initializers")
    jax.nn = nn
    jax.random = _Stub("jax.random")
    jax.config = _Stub("jax.config")
    jax.debug = _Stub("jax.debug")
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": linalg, "jax.lax": lax, "jax.nn": nn,
            "jax.nn.initializers": nn.initializers, "jax.tree_util": tree_util, "jax.random": jax.random}
    for name in ("jaxtyping", "flax", "flax.linen", "flax.linen.dtypes", "flax.linen.module", "flax.linen.normalization",
                 "flax.typing", "flax.core", "flax.training", "aqt", "aqt.jax", "aqt.jax.v2", "aqt.jax.v2.flax",
                 "aqt.jax.v2.flax.aqt_flax", "aqt.jax.v2.config", "aqt.jax.v2.aqt_dot_general"):
        mods[name] = _Stub(name)
    for name, mod in mods.items():                            # parent.child attribute for every dotted name
        if "." in name and name.rsplit(".", 1)[0] in mods and not name.startswith("jax"):
            setattr(mods[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], mod)
    mods["flax.linen"].Module = type("Module", (), {})       # a real base class: the reference subclasses it
    mods["flax.linen.module"].Module = mods["flax.linen"].Module
    sys.modules.update(mods)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return jax
