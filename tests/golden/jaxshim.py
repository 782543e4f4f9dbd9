"""NumPy-backed stand-ins for `jax`, `jax.numpy`, `jax.lax`, `jax.random`, `flax.struct`, `flax.core`, `chex`.

TEST INFRASTRUCTURE ONLY.  JAX is not installable in this image (no network), so the reference's
Overcooked environment cannot be imported as-is.  `install()` registers just enough of the JAX surface
in `sys.modules` for the UNMODIFIED reference source files

    jax_marl/environments/multi_agent_env.py
    jax_marl/environments/spaces.py
    jax_marl/environments/overcooked_environment/{common,layouts,overcooked}.py
    jax_marl/wrappers/baselines.py   (LogWrapper only)

to execute eagerly, line by line, on NumPy arrays.  `tests/golden/make_golden.py` uses this to run the
reference's own `reset / step / get_obs` and record golden trajectories.  What this does and does not pin:

  * the environment logic (movement, interact, pots, auto-reset, observation channels) is the
    reference's own code, executed;
  * `jax.random` is OUR restatement (`oracle/prng.py`) -- random draws are not pinned against real JAX;
  * dtypes follow NumPy promotion, which differs from JAX's (e.g. uint32 + int8 -> int64, not int32);
    values on this path never overflow either, and the golden files store each field cast to the dtype
    JAX would produce.
"""
import dataclasses
import sys
import types

import numpy as np

from oracle import prng as _prng


class _At:
    __slots__ = ("arr", "idx")

    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v, **_):
        out = np.array(self.arr, copy=True)
        out[self.idx] = np.asarray(v)  # jax casts the update to the operand dtype; so does numpy assignment
        return out.view(Arr)

    def get(self, **_):
        return _wrap(np.asarray(self.arr)[self.idx])

    def add(self, v, **_):
        out = np.array(self.arr, copy=True)
        np.add.at(out, self.idx, v)
        return out.view(Arr)


class _AtHelper:
    __slots__ = ("arr",)

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _At(self.arr, idx)


class Arr(np.ndarray):
    """ndarray with the functional-update `.at[...]` property."""

    @property
    def at(self):
        return _AtHelper(self)

    def __hash__(self):  # jnp arrays are unhashable too; only needed so dataclass eq/hash do not explode
        return id(self)


def _wrap(x):
    if isinstance(x, np.ndarray):
        return x.view(Arr)
    return x


def _w(fn):
    def inner(*a, **k):
        return _wrap(fn(*a, **k))
    inner.__name__ = getattr(fn, "__name__", "fn")
    return inner


def _array(obj, dtype=None, **_):
    if isinstance(obj, (list, tuple)):
        obj = [np.asarray(o) for o in obj] if len(obj) else obj
    return np.array(obj, dtype=dtype).view(Arr)


def _stack(arrs, axis=0, dtype=None):
    out = np.stack([np.asarray(a) for a in arrs], axis=axis)
    if dtype is not None:
        out = out.astype(dtype)  # jax converts freely (float32 -> uint8 at common.py:128-131)
    return out.view(Arr)


def _ones(shape, dtype=None):
    return np.ones(shape, dtype=np.float32 if dtype is None else dtype).view(Arr)


def _zeros(shape, dtype=None):
    return np.zeros(shape, dtype=np.float32 if dtype is None else dtype).view(Arr)


def _full(shape, fill, dtype=None):
    return np.full(shape, fill, dtype=dtype).view(Arr)


# ---------------------------------------------------------------------------------------------- pytrees

def _is_struct(x):
    return dataclasses.is_dataclass(x) and not isinstance(x, type)


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return type(tree)({k: tree_map(f, tree[k], *[r[k] for r in rest]) for k in tree})
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    if _is_struct(tree):
        kw = {fl.name: tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest])
              for fl in dataclasses.fields(tree)}
        return type(tree)(**kw)
    return f(tree, *rest)


def _struct_dataclass(cls):
    cls = dataclasses.dataclass(frozen=True, eq=False)(cls)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)
    cls.replace = replace
    return cls


# ---------------------------------------------------------------------------------------------- vmap

def _vmap(fn, in_axes=0, out_axes=0):
    """Eager loop over the mapped axis (axis 0 only, which is all the reference uses); arguments may be pytrees."""
    def leaves(t):
        out = []
        tree_map(lambda x: out.append(x), t)
        return out

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                assert ax == 0
                n = len(leaves(a)[0])
        outs = []
        for i in range(n):
            call = [tree_map(lambda x: _wrap(np.asarray(x)[i]), a) if ax is not None else a for a, ax in zip(args, axes)]
            outs.append(fn(*call))
        return tree_map(lambda *xs: _stack(xs), outs[0], *outs[1:])
    return mapped


def _scan(f, init, xs=None, length=None):
    """jax.lax.scan, eagerly."""
    if xs is None:
        n = length
    else:
        first = []
        tree_map(lambda a: first.append(a), xs)
        n = len(first[0])
    carry, ys = init, []
    for i in range(n):
        x = None if xs is None else tree_map(lambda a: _wrap(np.asarray(a)[i]), xs)
        carry, y = f(carry, x)
        ys.append(y)
    return carry, tree_map(lambda *v: _stack(v), ys[0], *ys[1:])


class _Config:
    @staticmethod
    def update(name, value):
        if name == "jax_threefry_partitionable":
            _Cfg.partitionable = bool(value)


def _jit(fn=None, static_argnums=None, **_):
    if fn is None:
        return lambda f: f
    return fn


# ---------------------------------------------------------------------------------------------- random

class _Cfg:
    partitionable = False


def _r_split(key, num=2):
    return _prng.split(np.asarray(key), num, _Cfg.partitionable).view(Arr)


def _r_randint(key, shape, minval, maxval, dtype=np.int32):
    n = int(np.prod(shape)) if len(shape) else 1
    out = _prng.randint(np.asarray(key), n, int(minval), int(maxval), _Cfg.partitionable)
    return out.reshape(shape).astype(dtype).view(Arr)


def _r_choice(key, a, shape=(), replace=True, p=None, axis=0):
    arr = np.asarray(a)
    n_inputs = arr.shape[0]
    n_draws = int(np.prod(shape)) if len(shape) else 1
    if p is None:
        assert replace, "shim: choice(replace=False, p=None) is not used by the reference path"
        ind = _prng.randint(np.asarray(key), n_draws, 0, n_inputs, _Cfg.partitionable)
    else:
        assert not replace, "shim: choice(replace=True, p=...) is not used by the reference path"
        ind = _prng.choice_no_replace_p(np.asarray(key), n_inputs, n_draws, np.asarray(p), _Cfg.partitionable)
    return arr[ind].reshape(shape).view(Arr)


# ---------------------------------------------------------------------------------------------- install

class _LinenModule:
    """flax.linen.Module stand-in: annotated class attributes become constructor fields; `apply(variables, x)`
    runs `__call__` with the submodules drawing their parameters from variables["params"] in creation order
    (`Dense_0`, `Dense_1`, ... -- flax's auto-naming for unnamed submodules inside an `nn.compact` method)."""
    _ctx = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = list(getattr(cls, "__annotations__", {}))

        def __init__(self, *a, **k):
            for name, val in zip(fields, a):
                setattr(self, name, val)
            for name in fields[len(a):]:
                setattr(self, name, k[name] if name in k else getattr(cls, name))
        cls.__init__ = __init__

    def apply(self, variables, *args):
        prev = _LinenModule._ctx
        _LinenModule._ctx = {"params": variables["params"], "count": {}}
        try:
            return self(*args)
        finally:
            _LinenModule._ctx = prev


class _LinenDense:
    def __init__(self, features, kernel_init=None, bias_init=None, use_bias=True, **_):
        self.features, self.use_bias = features, use_bias

    def __call__(self, x):
        ctx = _LinenModule._ctx
        i = ctx["count"].get("Dense", 0)
        ctx["count"]["Dense"] = i + 1
        p = ctx["params"]["Dense_%d" % i]
        kernel = np.asarray(p["kernel"], dtype=np.float32)
        assert kernel.shape[1] == self.features, (kernel.shape, self.features)
        y = np.asarray(x, dtype=np.float32) @ kernel          # promote_dtype(inputs, kernel): uint8 obs -> float32
        if self.use_bias:
            y = y + np.asarray(p["bias"], dtype=np.float32)
        return _wrap(y.astype(np.float32))


class _Categorical:
    def __init__(self, logits=None, probs=None):
        self.logits = logits


def install(partitionable=False):
    """Register the stand-in modules. Idempotent."""
    _Cfg.partitionable = bool(partitionable)
    if "jax" in sys.modules and getattr(sys.modules["jax"], "__is_oc_shim__", False):
        return
    assert "jax" not in sys.modules, "a real jax is importable; use it instead of the shim"

    jnp = types.ModuleType("jax.numpy")
    for name in ("uint8", "uint32", "int8", "int32", "float32", "bool_", "int_", "ndarray", "newaxis", "dtype",
                 "iinfo", "finfo", "inf", "pi"):
        setattr(jnp, name, getattr(np, name))
    jnp.array = _array
    jnp.asarray = _array
    jnp.stack = _stack
    jnp.ones = _ones
    jnp.zeros = _zeros
    jnp.full = _full
    for name in ("zeros_like", "ones_like", "arange", "where", "minimum", "maximum", "logical_and",
                 "logical_or", "logical_not", "any", "all", "sum", "expand_dims", "transpose", "tile",
                 "concatenate", "take", "squeeze", "reshape", "clip", "abs", "argmax", "cumsum", "mean"):
        setattr(jnp, name, _w(getattr(np, name)))

    lax = types.ModuleType("jax.lax")
    lax.select = lambda pred, a, b: _wrap(np.where(np.asarray(pred), a, b))
    lax.stop_gradient = lambda x: x
    lax.scan = _scan

    rnd = types.ModuleType("jax.random")
    rnd.PRNGKey = lambda seed: _prng.PRNGKey(seed).view(Arr)
    rnd.split = _r_split
    rnd.randint = _r_randint
    rnd.choice = _r_choice

    tree = types.ModuleType("jax.tree")
    tree.map = tree_map

    jax = types.ModuleType("jax")
    jax.__is_oc_shim__ = True
    jax.__path__ = []
    jax.numpy, jax.lax, jax.random, jax.tree = jnp, lax, rnd, tree
    jax.vmap = _vmap
    jax.jit = _jit
    jax.config = _Config
    jax.block_until_ready = lambda x: x
    jax.tree_map = tree_map

    chex = types.ModuleType("chex")
    chex.Array = np.ndarray
    chex.PRNGKey = np.ndarray

    flax = types.ModuleType("flax")
    flax.__path__ = []
    struct = types.ModuleType("flax.struct")
    struct.dataclass = _struct_dataclass
    core = types.ModuleType("flax.core")
    core.__path__ = []
    fd = types.ModuleType("flax.core.frozen_dict")

    class FrozenDict(dict):
        def __hash__(self):
            return id(self)
    fd.FrozenDict = FrozenDict
    fd.freeze = lambda d: FrozenDict(d)
    fd.unfreeze = lambda d: dict(d)
    core.frozen_dict = fd
    core.FrozenDict = FrozenDict
    flax.struct, flax.core = struct, core
    trav = types.ModuleType("flax.traverse_util")
    trav.flatten_dict = trav.unflatten_dict = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    flax.traverse_util = trav

    # only what `jax_marl/wrappers/baselines.py` imports at module level
    gymnax = types.ModuleType("gymnax"); gymnax.__path__ = []
    genv = types.ModuleType("gymnax.environments"); genv.__path__ = []
    gsp = types.ModuleType("gymnax.environments.spaces")
    gsp.Box = gsp.Discrete = type("_Space", (), {})
    saf = types.ModuleType("safetensors"); saf.__path__ = []
    saff = types.ModuleType("safetensors.flax")
    saff.save_file = saff.load_file = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())

    # flax.linen / distrax: just enough to execute architectures/mlp.py (ActorCritic.__call__) unmodified
    linen = types.ModuleType("flax.linen")
    linen.__path__ = []
    linen.Module = _LinenModule
    linen.Dense = _LinenDense
    linen.compact = lambda f: f
    linen.relu = lambda x: _wrap(np.maximum(np.asarray(x), np.float32(0)))
    linen.tanh = lambda x: _wrap(np.tanh(np.asarray(x, dtype=np.float32)))
    inits = types.ModuleType("flax.linen.initializers")
    inits.constant = lambda v: ("constant", v)
    inits.orthogonal = lambda scale=1.0: ("orthogonal", scale)
    linen.initializers = inits
    flax.linen = linen
    distrax = types.ModuleType("distrax")
    distrax.Categorical = _Categorical

    sys.modules.update({
        "flax.linen": linen, "flax.linen.initializers": inits, "distrax": distrax,
        "jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.random": rnd, "jax.tree": tree,
        "chex": chex, "flax": flax, "flax.struct": struct, "flax.core": core,
        "flax.core.frozen_dict": fd, "flax.traverse_util": trav,
        "gymnax": gymnax, "gymnax.environments": genv, "gymnax.environments.spaces": gsp,
        "safetensors": saf, "safetensors.flax": saff,
    })


def load_reference(ref_root="/root/reference", partitionable=False):
    """Import the reference's hot-path modules from `ref_root` by file path, bypassing the package
    `__init__`s that would drag in the MPE / STORM environments.  Returns a namespace with
    `Overcooked`, `State`, `layouts` (dict name -> FrozenDict), `LogWrapper`, `common`."""
    import importlib.util
    import os

    install(partitionable)

    def _load(modname, relpath):
        if modname in sys.modules:
            return sys.modules[modname]
        spec = importlib.util.spec_from_file_location(modname, os.path.join(ref_root, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    def _pkg(name):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        return sys.modules[name]

    _pkg("jax_marl")
    envs = _pkg("jax_marl.environments")
    mae = _load("jax_marl.environments.multi_agent_env", "jax_marl/environments/multi_agent_env.py")
    envs.MultiAgentEnv, envs.State = mae.MultiAgentEnv, mae.State
    envs.multi_agent_env = mae
    envs.spaces = _load("jax_marl.environments.spaces", "jax_marl/environments/spaces.py")
    oe = _pkg("jax_marl.environments.overcooked_environment")
    base = "jax_marl/environments/overcooked_environment/"
    oe.common = _load("jax_marl.environments.overcooked_environment.common", base + "common.py")
    oe.layouts = _load("jax_marl.environments.overcooked_environment.layouts", base + "layouts.py")
    oe.overcooked = _load("jax_marl.environments.overcooked_environment.overcooked", base + "overcooked.py")
    _pkg("jax_marl.wrappers")
    wb = _load("jax_marl.wrappers.baselines", "jax_marl/wrappers/baselines.py")

    ns = types.SimpleNamespace()
    ns.Overcooked = oe.overcooked.Overcooked
    ns.State = oe.overcooked.State
    ns.layouts = oe.layouts.overcooked_layouts
    ns.layouts_module = oe.layouts
    ns.common = oe.common
    ns.LogWrapper = wb.LogWrapper
    ns.FrozenDict = sys.modules["flax.core.frozen_dict"].FrozenDict
    ns.ActorCritic = _load("architectures.mlp", "architectures/mlp.py").ActorCritic
    return ns
