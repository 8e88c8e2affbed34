"""numpy-backed stand-ins for the parts of jax / flax the reference's NeRF-SH forward path imports, so that
/root/reference/nerf_sh/nerf/model_utils.py and models.py (NerfModel.__call__) can be EXECUTED here, unmodified,
to generate golden vectors (jax, jaxlib and flax are not installable in this image).

What is faithful: the reference's own source text runs — every formula, axis, concatenation order, clamp and
branch — on float32 numpy arrays (jnp's default dtype policy is mimicked: float64 results are cast to float32).
What is not: XLA's arithmetic (its exp / sin polynomials and reduction orders differ from numpy's in the last
bits) and jax.random's threefry streams — random draws are INJECTED through the key objects instead.

Only tests/golden/make_golden.py imports this; nothing in the product or in the GPU tests does.
"""
import sys
import types

import numpy as np


# ---------------------------------------------------------------- jax.numpy -----------------------------------
def _f32(x):
    if isinstance(x, np.ndarray) and x.dtype == np.float64:
        return x.astype(np.float32)
    if isinstance(x, np.float64):
        return np.float32(x)
    if isinstance(x, tuple):
        return tuple(_f32(v) for v in x)
    return x


def _wrap(fn):
    def g(*a, **k):
        return _f32(fn(*a, **k))
    g.__name__ = getattr(fn, "__name__", "wrapped")
    return g


def _make_jnp():
    jnp = types.ModuleType("jax.numpy")
    for name in ("linspace", "concatenate", "broadcast_to", "array", "reshape", "sin", "cos", "stack", "exp",
                 "ones_like", "zeros_like", "cumprod", "cumsum", "where", "sum", "maximum", "minimum", "zeros", "ones",
                 "max", "min", "clip", "nan_to_num", "sort", "tile", "prod", "mean", "sqrt", "log", "abs", "arange",
                 "expand_dims", "squeeze", "transpose", "matmul", "dot", "square", "power", "sign"):
        setattr(jnp, name, _wrap(getattr(np, name)))
    # jnp reductions accept a list of axes; numpy wants a tuple
    jnp.mean = _wrap(lambda x, axis=None, **k: np.mean(x, axis=tuple(axis) if isinstance(axis, list) else axis, **k))
    jnp.pi = np.pi
    jnp.float32 = np.float32
    jnp.finfo = np.finfo
    jnp.ndarray = np.ndarray
    jnp.newaxis = None
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.norm = _wrap(np.linalg.norm)
    jnp.linalg = linalg
    return jnp


# ---------------------------------------------------------------- jax.random ----------------------------------
class Key:
    """Stand-in for a PRNG key: carries the arrays the test wants `uniform` / `normal` to return (matched by shape);
    children of `split` share them.  Unmatched draws come from a seeded numpy RandomState."""

    def __init__(self, uniform=None, normal=None, seed=0):
        self.uniform = uniform
        self.normal = normal
        self.seed = seed


def _make_random():
    random = types.ModuleType("jax.random")
    random.PRNGKey = lambda seed: Key(seed=int(seed))

    def split(key, num=2):
        return [Key(key.uniform, key.normal, seed=key.seed * 7919 + i + 1) for i in range(num)]

    def uniform(key, shape, dtype=np.float32, minval=0.0, maxval=1.0):
        shape = tuple(shape)
        if key.uniform is not None and tuple(key.uniform.shape) == shape:
            u = np.asarray(key.uniform, dtype=np.float32)
        else:
            u = np.random.RandomState(key.seed % (2 ** 31)).uniform(size=shape).astype(np.float32)
        return (u * np.float32(maxval - minval) + np.float32(minval)).astype(np.float32)

    def normal(key, shape, dtype=np.float32):
        shape = tuple(shape)
        if key.normal is not None and tuple(key.normal.shape) == shape:
            return np.asarray(key.normal, dtype=np.float32)
        return np.random.RandomState(key.seed % (2 ** 31)).normal(size=shape).astype(np.float32)

    random.split, random.uniform, random.normal = split, uniform, normal
    return random


# ---------------------------------------------------------------- flax.linen ----------------------------------
_PARAM_STACK = []


class Module:
    """dataclass-like construction from keyword arguments (class attributes are the defaults), then setup()."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        if hasattr(self, "setup"):
            self.setup()

    def variable(self, col, name, init, *args):
        return types.SimpleNamespace(value=init(*args))

    def apply(self, variables, *args, method=None, **kw):
        """model.apply(variables, ...): bind {"params": {"MLP_i": {"Dense_j": {"kernel", "bias"}}}} to the
        sub-modules (flax's own tree layout, octree/nerf/models.py:75-87), then call `method` or __call__."""
        for name, layers in variables["params"].items():
            sub = getattr(self, name)
            sub._params = [(layers[f"Dense_{j}"]["kernel"], layers[f"Dense_{j}"]["bias"]) for j in range(len(layers))]
        return (method or self.__call__)(*args, **kw)


def compact(fn):
    """@nn.compact: while the method runs, nn.Dense layers take their (kernel, bias) in creation order —
    Dense_0, Dense_1, ... exactly as flax names them — from the module's `_params` list."""
    def g(self, *a, **k):
        _PARAM_STACK.append(iter(self._params))
        try:
            return fn(self, *a, **k)
        finally:
            _PARAM_STACK.pop()
    return g


class Dense:
    def __init__(self, features, kernel_init=None, **kw):
        self.features = features

    def __call__(self, x):
        w, b = next(_PARAM_STACK[-1])
        w, b = np.asarray(w, dtype=np.float32), np.asarray(b, dtype=np.float32)
        assert w.shape == (x.shape[-1], self.features), (w.shape, x.shape, self.features)
        return (x.astype(np.float32) @ w + b).astype(np.float32)


def _make_linen():
    nn = types.ModuleType("flax.linen")
    nn.Module = Module
    nn.compact = compact
    nn.Dense = Dense
    nn.relu = lambda x: np.maximum(x, np.float32(0)).astype(np.float32)
    nn.sigmoid = lambda x: (np.float32(1) / (np.float32(1) + np.exp(-x.astype(np.float32)))).astype(np.float32)
    nn.softplus = lambda x: np.logaddexp(x, 0).astype(np.float32)
    return nn


def _tree_leaves(t):
    if isinstance(t, dict):
        for k in sorted(t):
            yield from _tree_leaves(t[k])
    else:
        yield t


def _tree_map(fn, t):
    return {k: _tree_map(fn, v) for k, v in t.items()} if isinstance(t, dict) else fn(t)


def install():
    """register the stand-ins in sys.modules; returns the names so that the caller can remove them again."""
    jax = types.ModuleType("jax")
    jnp = _make_jnp()
    random = _make_random()
    lax = types.ModuleType("jax.lax")
    lax.stop_gradient = lambda x: x
    jnn = types.ModuleType("jax.nn")
    jnn.initializers = types.SimpleNamespace(glorot_uniform=lambda: None)
    jnn.relu = lambda x: np.maximum(x, np.float32(0)).astype(np.float32)
    lax.pmean = lambda x, axis_name=None: x                      # one device
    jax.numpy, jax.random, jax.lax, jax.nn = jnp, random, lax, jnn
    # jax.value_and_grad: the VALUE side runs the reference's loss_fn; gradients cannot be taken of numpy code and
    # are returned as zeros (the oracle's gradient is autograd of its pinned forward)
    def value_and_grad(fn, has_aux=False):
        def g(x):
            return fn(x), _tree_map(lambda a: np.zeros_like(a), x)
        return g
    jax.value_and_grad = value_and_grad
    tree_util = types.ModuleType("jax.tree_util")
    def tree_reduce(f, tree, initializer=None):
        acc = initializer
        for leaf in _tree_leaves(tree):
            acc = f(acc, leaf)
        return acc
    tree_util.tree_reduce = tree_reduce
    jax.tree_util = tree_util
    jax.config = types.SimpleNamespace(parse_flags_with_absl=lambda: None)
    cfg_mod = types.ModuleType("jax.config")
    jax.dlpack = types.ModuleType("jax.dlpack")
    jax.scipy = types.ModuleType("jax.scipy")
    import scipy.signal as _ss
    jax.scipy.signal = types.SimpleNamespace(convolve2d=_wrap(lambda a, b, mode="full": _ss.convolve2d(a, b, mode=mode)))

    def vmap(fn, in_axes=0, out_axes=0):
        """jax.vmap for one array argument: apply fn to every slice along in_axes, stack along out_axes."""
        def g(x):
            xs = np.moveaxis(np.asarray(x), in_axes, 0)
            return np.moveaxis(np.stack([fn(v) for v in xs], axis=0), 0, out_axes)
        return g
    jax.vmap = vmap
    jax.host_id = lambda: 0
    jax.host_count = lambda: 1
    jax.local_device_count = lambda: 1
    jax.device_count = lambda: 1
    flax = types.ModuleType("flax")
    linen = _make_linen()
    flax.linen = linen
    import dataclasses
    flax.struct = types.SimpleNamespace(dataclass=lambda cls: _with_replace(dataclasses.dataclass(cls)))
    flax.optim = types.SimpleNamespace(Optimizer=object)
    metrics = types.ModuleType("flax.metrics")
    metrics.tensorboard = types.ModuleType("flax.metrics.tensorboard")
    training = types.ModuleType("flax.training")
    training.checkpoints = types.ModuleType("flax.training.checkpoints")
    flax.metrics, flax.training = metrics, training
    mods = {"jax": jax, "jax.numpy": jnp, "jax.random": random, "jax.lax": lax, "jax.nn": jnn, "jax.tree_util": tree_util,
            "jax.dlpack": jax.dlpack, "jax.scipy": jax.scipy, "flax": flax, "flax.linen": linen, "flax.metrics": metrics,
            "flax.metrics.tensorboard": metrics.tensorboard, "flax.training": training,
            "flax.training.checkpoints": training.checkpoints}
    sys.modules.update(mods)
    return list(mods)


def _with_replace(cls):
    import dataclasses
    cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
    return cls


def uninstall(names):
    for n in names:
        sys.modules.pop(n, None)
