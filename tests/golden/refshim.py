"""NumPy stand-ins for the third-party packages the reference imports but this image lacks
(jax 0.4.13, diffrax 0.4.0, tensorflow-probability 0.20.1, jaxtyping, optax, blackjax, fastprogress, matplotlib ...).

Purpose: let ``make_golden.py`` import and run the reference's OWN hot-path files, unmodified, from
``/root/reference`` so that golden vectors pin the reference's orchestration (step order, indexing, the
``trace(H_t @ P)`` quirk, psd_solve boost, symmetrize placement ...).  The arithmetic that lives inside the absent
packages is restated here independently of ``oracle/cd_oracle.py`` (LAPACK Cholesky / triangular solves instead of the
oracle's hand-rolled loops; per-trajectory Python scalars instead of batched masks), so the two restatements check
each other.  Used only when generating fixtures in the build container; nothing here runs on the GPU box.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import math
import sys
import types

import numpy as np
import scipy.linalg as sla

# ---------------------------------------------------------------------------------------------------------------------
# permissive auto-mock for everything we do not implement
# ---------------------------------------------------------------------------------------------------------------------


class _MockMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _make_mock(f"{cls.__name__}.{name}")

    def __getitem__(cls, item):
        return cls

    def __or__(cls, other):
        return cls

    def __ror__(cls, other):
        return cls


def _make_mock(name):
    return _MockMeta(name, (_MockBase,), {})


class _MockBase(metaclass=_MockMeta):
    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__()

    def __call__(self, *a, **k):
        # decorator use: @mock(...) def f -> return f
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return self

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _make_mock(name)()

    def __getitem__(self, item):
        return self

    def __iter__(self):
        return iter(())


class _AutoModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        return _make_mock(name)


_MOCK_ROOTS = (
    "jax", "jaxlib", "jaxtyping", "diffrax", "equinox", "tensorflow_probability", "optax", "blackjax", "fastprogress",
    "matplotlib", "seaborn", "sklearn_unused",
)


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _MOCK_ROOTS and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _AutoModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ---------------------------------------------------------------------------------------------------------------------
# jax.numpy  (numpy, float64, NaN-propagating Cholesky)
# ---------------------------------------------------------------------------------------------------------------------
_DTYPE = [np.float64]


def set_default_dtype(dt):
    _DTYPE[0] = np.dtype(dt).type


def _chol(A):
    """jnp.linalg.cholesky: jax 0.4.13 forwards to lax.linalg.cholesky(x, symmetrize_input=True), i.e. it factors
    (A + A^T) / 2.  (Exact no-op for a symmetric A; it matters where the reference hands it a non-symmetric matrix: the
    1-D `emissions.cov` broadcast in cd_linear/inference.py:613.)"""
    A = np.asarray(A)
    if not np.all(np.isfinite(A)):
        return np.full_like(A, np.nan)
    A = 0.5 * (A + np.swapaxes(A, -1, -2))
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        return np.full_like(A, np.nan)


class _JnpLinalg(types.ModuleType):
    def __getattr__(self, name):
        return getattr(np.linalg, name)


_jnp_linalg = _JnpLinalg("jax.numpy.linalg")
_jnp_linalg.cholesky = _chol


def _asf(x, dtype=None):
    a = np.asarray(x)
    if dtype is not None:
        return a.astype(dtype)
    if a.dtype.kind == "f" and a.dtype != _DTYPE[0]:
        return a.astype(_DTYPE[0])
    return a


class _Jnp(types.ModuleType):
    def __getattr__(self, name):
        return getattr(np, name)


jnp = _Jnp("jax.numpy")
jnp.linalg = _jnp_linalg
jnp.ndarray = np.ndarray
jnp.array = lambda x, dtype=None: _asf(x, dtype)
jnp.asarray = lambda x, dtype=None: _asf(x, dtype)
jnp.zeros = lambda shape, dtype=None: np.zeros(shape, dtype or _DTYPE[0])
jnp.ones = lambda shape, dtype=None: np.ones(shape, dtype or _DTYPE[0])
jnp.eye = lambda n, m=None, dtype=None: np.eye(n, m, dtype=dtype or _DTYPE[0])
jnp.float32 = np.float32
jnp.float64 = np.float64


# ---------------------------------------------------------------------------------------------------------------------
# pytrees, scan, vmap, jit, jacobians
# ---------------------------------------------------------------------------------------------------------------------
def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_map(f, tree, *rest):
    if tree is None:
        return None
    if _is_namedtuple(tree):
        return type(tree)(*(tree_map(f, c, *(r[i] for r in rest)) for i, c in enumerate(tree)))
    if isinstance(tree, (tuple, list)):
        return type(tree)(tree_map(f, c, *(r[i] for r in rest)) for i, c in enumerate(tree))
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *(r[k] for r in rest)) for k, v in tree.items()}
    return f(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out


def _index_tree(tree, i):
    return tree_map(lambda x: x[i], tree)


def _stack_trees(trees):
    first = trees[0]
    if first is None:
        return None
    return tree_map(lambda *xs: np.stack([np.asarray(x) for x in xs], axis=0), first, *trees[1:])


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = length
    else:
        n = len(tree_leaves(xs)[0])
    idx = range(n - 1, -1, -1) if reverse else range(n)
    carry = init
    ys = [None] * n
    for i in idx:
        x_i = _index_tree(xs, i) if xs is not None else None
        carry, y = f(carry, x_i)
        ys[i] = y
    return carry, (_stack_trees(ys) if n > 0 else None)


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = len(tree_leaves(a)[0]) if ax == 0 else np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            call = []
            for a, ax in zip(args, axes):
                if ax is None:
                    call.append(a)
                elif ax == 0:
                    call.append(_index_tree(a, i))
                else:
                    call.append(np.take(a, i, axis=ax))
            outs.append(fun(*call))
        return _stack_trees(outs)

    return mapped


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


class _JacFn:
    """Jacobian w.r.t. argument 0 by complex-step differentiation (exact to rounding for analytic f)."""

    def __init__(self, f):
        self.f = f

    def __call__(self, x, *args):
        x = np.asarray(x, np.float64)
        h = 1e-30
        cols = []
        for k in range(x.shape[0]):
            xc = x.astype(np.complex128)
            xc[k] += 1j * h
            cols.append(np.imag(np.asarray(self.f(xc, *args))) / h)
        return np.stack(cols, axis=-1).astype(_DTYPE[0])


class _HessFn:
    """jacfwd(jacrev(f)): central difference (h = 1) of the complex-step Jacobian.
    Exact (to rounding) for drifts that are polynomials of degree <= 2 -- all registry drifts are."""

    def __init__(self, jf):
        self.jf = jf

    def __call__(self, x, *args):
        x = np.asarray(x, np.float64)
        cols = []
        for k in range(x.shape[0]):
            e = np.zeros_like(x)
            e[k] = 1.0
            cols.append((self.jf(x + e, *args).astype(np.float64) - self.jf(x - e, *args).astype(np.float64)) / 2.0)
        return np.stack(cols, axis=-1).astype(_DTYPE[0])


def jacfwd(f, argnums=0):
    if isinstance(f, _JacFn):
        return _HessFn(f)
    return _JacFn(f)


jacrev = jacfwd


# ---------------------------------------------------------------------------------------------------------------------
# jax.random (NOT threefry: EnKF/sampling goldens are only distribution-level)
# ---------------------------------------------------------------------------------------------------------------------
def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], np.uint32)


def _rng(key):
    return np.random.default_rng(np.random.SeedSequence([int(v) for v in np.asarray(key).ravel()]))


def split(key, num=2):
    return _rng(key).integers(0, 2**32, size=(num, 2), dtype=np.uint32)


def normal(key, shape=(), dtype=None):
    return _rng(key).standard_normal(shape).astype(_DTYPE[0])


def multivariate_normal(key, mean, cov, shape=None, dtype=None, method="cholesky"):
    mean, cov = np.asarray(mean), np.asarray(cov)
    shape = tuple(shape or ())
    z = _rng(key).standard_normal(shape + mean.shape[-1:])
    return (mean + z @ _chol(cov).T).astype(_DTYPE[0])


# ---------------------------------------------------------------------------------------------------------------------
# jax.scipy.linalg
# ---------------------------------------------------------------------------------------------------------------------
def cho_factor(A, lower=False):
    L = _chol(A)
    return (L if lower else L.T), lower


def cho_solve(c_and_lower, b):
    c, lower = c_and_lower
    if not np.all(np.isfinite(c)):
        return np.full(np.broadcast_shapes(np.shape(b)), np.nan)
    L = c if lower else c.T
    y = sla.solve_triangular(L, b, lower=True)
    return sla.solve_triangular(L.T, y, lower=False)


# ---------------------------------------------------------------------------------------------------------------------
# TFP MultivariateNormalFullCovariance
# ---------------------------------------------------------------------------------------------------------------------
class MultivariateNormalFullCovariance:
    def __init__(self, loc=None, covariance_matrix=None, **kw):
        self.loc = np.asarray(loc)
        self.cov = np.asarray(covariance_matrix)

    def log_prob(self, y):
        L = _chol(self.cov)
        m = self.cov.shape[-1]
        r = np.atleast_1d(np.asarray(y) - self.loc)
        if not np.all(np.isfinite(L)):
            return _DTYPE[0](np.nan)
        z = sla.solve_triangular(L, r, lower=True)
        return _DTYPE[0](-0.5 * float(z @ z) - float(np.sum(np.log(np.diag(L)))) - 0.5 * m * math.log(2.0 * math.pi))

    def mean(self):
        return self.loc

    def covariance(self):
        return self.cov


# ---------------------------------------------------------------------------------------------------------------------
# diffrax: fixed-step explicit RK under ConstantStepSize, SaveAt(t1=True)
# ---------------------------------------------------------------------------------------------------------------------
class _Solver:
    a = None
    b = None

    def __init__(self, *a, **k):
        pass


class Euler(_Solver):
    a, b = [[]], [1.0]


class Heun(_Solver):
    a, b = [[], [1.0]], [0.5, 0.5]


class Midpoint(_Solver):
    a, b = [[], [0.5]], [0.0, 1.0]


class Ralston(_Solver):
    a, b = [[], [0.75]], [1 / 3, 2 / 3]


class Bosh3(_Solver):
    a, b = [[], [0.5], [0.0, 0.75]], [2 / 9, 1 / 3, 4 / 9]


class Rk4(_Solver):  # shim-only extension: diffrax 0.4.0 has no classical RK4
    a, b = [[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]], [1 / 6, 1 / 3, 1 / 3, 1 / 6]


class Dopri5(_Solver):
    a = [
        [],
        [1 / 5],
        [3 / 40, 9 / 40],
        [44 / 45, -56 / 15, 32 / 9],
        [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
        [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    ]
    b = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]


class ConstantStepSize:
    def __init__(self, *a, **k):
        pass


class RecursiveCheckpointAdjoint:
    def __init__(self, *a, **k):
        pass


class DirectAdjoint(RecursiveCheckpointAdjoint):
    pass


class SaveAt:
    def __init__(self, t1=False, **k):
        assert t1


class ODETerm:
    def __init__(self, vf):
        self.vf = vf


class ControlTerm:
    def __init__(self, vf, control):
        self.vf, self.control = vf, control


class MultiTerm:
    def __init__(self, *terms):
        self.terms = terms


class VirtualBrownianTree:
    def __init__(self, t0, t1, tol, shape, key):
        self.shape, self.key, self.count = shape, key, 0

    def increment(self, dt):
        self.count += 1
        k = np.concatenate([np.asarray(self.key).ravel(), [self.count]])
        return np.sqrt(dt) * np.random.default_rng(np.random.SeedSequence([int(v) for v in k])).standard_normal(self.shape)


class _Sol:
    def __init__(self, ys):
        self.ys = ys


def _tm2(f, a, b):
    return tree_map(f, a, b)


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, saveat=None, stepsize_controller=None, adjoint=None,
                max_steps=4096, **kw):
    assert isinstance(stepsize_controller, ConstantStepSize), "only ConstantStepSize is restated"
    sde = isinstance(terms, MultiTerm)
    drift = terms.terms[0].vf if sde else terms.vf
    dtype = np.asarray(tree_leaves(y0)[0]).dtype
    dtype = dtype.type if dtype.kind == "f" else _DTYPE[0]
    t0, t1, dt0 = dtype(t0), dtype(t1), dtype(dt0)
    tol = dtype(1e-10 if dtype is np.float64 else 1e-6)
    y = tree_map(lambda c: np.asarray(c, dtype), y0)
    tprev, tnext = t0, min(t0 + dt0, t1)
    n = 0
    a, b = solver.a, solver.b
    while tprev < t1:
        if n >= max_steps:
            y = tree_map(lambda c: np.full_like(c, np.nan), y)
            break
        dt = tnext - tprev
        noise = None
        if sde:
            dW = terms.terms[1].control.increment(dt)
        ks = []
        for i in range(len(b)):
            yi = y
            for j in range(i):
                if a[i][j] != 0.0:
                    yi = _tm2(lambda c, k, aij=dtype(a[i][j]): c + aij * k, yi, ks[j])
            ti = tprev + dtype(sum(a[i])) * dt
            k = tree_map(lambda f: dt * np.asarray(f, dtype), drift(ti, yi, args))
            if sde:
                g = np.asarray(terms.terms[1].vf(ti, yi, args), dtype)
                k = k + (g @ dW if g.ndim == 2 else g * dW)
            ks.append(k)
        for j in range(len(b)):
            if b[j] != 0.0:
                y = _tm2(lambda c, k, bj=dtype(b[j]): c + bj * k, y, ks[j])
        n += 1
        tprev = tnext
        cand = tprev + dt0
        tnext = t1 if cand > t1 - tol else cand
    return _Sol(tree_map(lambda c: np.asarray(c)[None, ...], y))


# ---------------------------------------------------------------------------------------------------------------------
# install
# ---------------------------------------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = _AutoModule(name)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(reference_root="/root/reference"):
    if "jax" in sys.modules and not isinstance(sys.modules["jax"], _AutoModule):
        raise RuntimeError("a real jax is importable: use it instead of the shims")
    sys.meta_path.insert(0, _MockFinder())
    lax = _mod("jax.lax", scan=scan)
    tree_util = _mod(
        "jax.tree_util", tree_map=tree_map, tree_leaves=tree_leaves, register_pytree_node_class=lambda c: c
    )
    jr = _mod("jax.random", PRNGKey=PRNGKey, split=split, normal=normal, multivariate_normal=multivariate_normal)
    jsl = _mod("jax.scipy.linalg", cho_factor=cho_factor, cho_solve=cho_solve)
    jsp = _mod("jax.scipy", linalg=jsl)
    sys.modules["jax.numpy"] = jnp
    sys.modules["jax.numpy.linalg"] = _jnp_linalg
    dbg = _mod("jax.debug")
    cfg = types.SimpleNamespace(update=lambda *a, **k: None)
    _mod(
        "jax", numpy=jnp, lax=lax, tree_util=tree_util, random=jr, scipy=jsp, debug=dbg, vmap=vmap, jit=jit,
        jacfwd=jacfwd, jacrev=jacrev, tree_map=tree_map, config=cfg,
    )
    dfx_attrs = {k: v for k, v in globals().items() if k in (
        "Euler", "Heun", "Midpoint", "Ralston", "Bosh3", "Rk4", "Dopri5", "ConstantStepSize",
        "RecursiveCheckpointAdjoint", "DirectAdjoint", "SaveAt", "ODETerm", "ControlTerm", "MultiTerm",
        "VirtualBrownianTree", "diffeqsolve")}
    _mod("diffrax", AbstractSolver=_Solver, AbstractStepSizeController=ConstantStepSize,
         AbstractAdjoint=RecursiveCheckpointAdjoint, **dfx_attrs)
    tfd = _mod("tensorflow_probability.substrates.jax.distributions",
               MultivariateNormalFullCovariance=MultivariateNormalFullCovariance)
    tfb = _mod("tensorflow_probability.substrates.jax.bijectors")
    sj = _mod("tensorflow_probability.substrates.jax", distributions=tfd, bijectors=tfb)
    sub = _mod("tensorflow_probability.substrates", jax=sj)
    _mod("tensorflow_probability", substrates=sub)
    for p in (f"{reference_root}/src", reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
