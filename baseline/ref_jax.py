"""Drives the REFERENCE's own Overcooked environment through its public API under `jax.jit(jax.vmap(...))`.

Not product code and not the oracle: this is the harness behind `bench.py --impl reference` (the reference's jitted, vmapped
step on JAX_PLATFORMS=cpu, timed on the host cores) and behind `tests/test_real_jax.py` (real JAX against the CPU oracle on
the same keys and actions).  It needs an importable `jax` and the reference tree (`baseline/_ref/` or `/root/reference`);
neither exists in the build image or on the GPU boxes, so every caller probes with `available()` and falls back (bench: the
C port, labelled as such; tests: skip).  The wiring itself is exercised in the build container by running this very module
on `tests/golden/jaxshim.py` installed as a stand-in `jax` (tests/test_real_jax.py::test_harness_wiring_on_the_shim).

What it mirrors: the callers' recipe of baselines/IPPO_CL.py:477-478, 535-538 and the random-action rollout of
baselines/unused/random_actions.py:197-225 --

    obsv, env_state = jax.vmap(env.reset, in_axes=(0,))(reset_rng)
    obsv, env_state, reward, done, info = jax.vmap(env.step, in_axes=(0, 0, 0))(step_rng, env_state, actions)

inside `jax.lax.scan`, one XLA program per layout.
"""
import importlib.util
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
LAYOUT_FIELDS = ("wall_idx", "agent_idx", "goal_idx", "plate_pile_idx", "onion_pile_idx", "pot_idx")
DYN = {"agent_pos": np.uint32, "agent_dir": np.int8, "agent_dir_idx": np.int32, "agent_inv": np.int32,
       "maze_map": np.uint8, "time": np.int32, "terminal": np.bool_}


def find_reference():
    """Directory that holds the reference's `jax_marl/` package, or None."""
    for p in REF_CANDIDATES:
        if os.path.isdir(os.path.join(p, "jax_marl", "environments", "overcooked_environment")):
            return p
    return None


def import_jax():
    """The real `jax` (CPU platform unless the caller chose one), or None.  A stand-in registered by the test shim counts as
    absent here; tests that want the shim pass the module explicitly."""
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    try:
        import jax
    except Exception:
        return None
    if getattr(jax, "__is_oc_shim__", False):
        return None
    return jax


def available():
    return import_jax() is not None and find_reference() is not None


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    return sys.modules[name]


def load_reference(ref_root=None):
    """Import the reference's hot-path modules BY FILE PATH (the package `__init__`s would drag in the MPE / STORM environments
    and their dependencies).  Returns a namespace: Overcooked, LogWrapper, FrozenDict."""
    ref_root = ref_root or find_reference()
    if ref_root is None:
        raise RuntimeError("reference tree not found (baseline/_ref or /root/reference)")

    def _load(modname, relpath):
        if modname in sys.modules:
            return sys.modules[modname]
        spec = importlib.util.spec_from_file_location(modname, os.path.join(ref_root, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    # optional third-party modules that wrappers/baselines.py imports at module level but the LogWrapper never touches
    for opt in ("gymnax", "safetensors"):
        try:
            __import__(opt)
        except Exception:
            if opt == "gymnax":
                _stub("gymnax"); _stub("gymnax.environments")
                sp = _stub("gymnax.environments.spaces")
                sp.Box = sp.Discrete = type("_Space", (), {})
            else:
                _stub("safetensors")
                sf = _stub("safetensors.flax")
                sf.save_file = sf.load_file = None
    _stub("jax_marl")
    envs = _stub("jax_marl.environments")
    mae = _load("jax_marl.environments.multi_agent_env", "jax_marl/environments/multi_agent_env.py")
    envs.MultiAgentEnv, envs.State, envs.multi_agent_env = mae.MultiAgentEnv, mae.State, mae
    envs.spaces = _load("jax_marl.environments.spaces", "jax_marl/environments/spaces.py")
    oe = _stub("jax_marl.environments.overcooked_environment")
    base = "jax_marl/environments/overcooked_environment/"
    oe.common = _load("jax_marl.environments.overcooked_environment.common", base + "common.py")
    oe.layouts = _load("jax_marl.environments.overcooked_environment.layouts", base + "layouts.py")
    oe.overcooked = _load("jax_marl.environments.overcooked_environment.overcooked", base + "overcooked.py")
    _stub("jax_marl.wrappers")
    wb = _load("jax_marl.wrappers.baselines", "jax_marl/wrappers/baselines.py")
    ns = types.SimpleNamespace()
    ns.Overcooked, ns.LogWrapper = oe.overcooked.Overcooked, wb.LogWrapper
    ns.FrozenDict = sys.modules["flax.core.frozen_dict"].FrozenDict
    return ns


def make_env(jax, ns, layout, name, *, random_reset=False, max_steps=400, task_id=0, log_wrapper=False):
    jnp = jax.numpy
    d = {"height": int(layout["height"]), "width": int(layout["width"])}
    for f in LAYOUT_FIELDS:
        d[f] = jnp.array(np.asarray(layout[f], dtype=np.int32))
    env = ns.Overcooked(layout=ns.FrozenDict(d), layout_name=name, random_reset=random_reset, max_steps=max_steps,
                        task_id=task_id)
    return ns.LogWrapper(env) if log_wrapper else env


def set_threefry_partitionable(jax, flag):
    jax.config.update("jax_threefry_partitionable", bool(flag))


def rollout(jax, ns, layout, name, *, random_reset, max_steps, reset_keys, step_keys, actions, task_id=0,
            partitionable=False):
    """reset + T vmapped steps on given raw keys (uint32 [E,2] / [T,E,2]) and actions (int32 [T,E,2]); returns the same
    record layout as tests/golden/*.npz (r_* after reset, s_* per step), each field cast to the dtype JAX documents."""
    jnp = jax.numpy
    set_threefry_partitionable(jax, partitionable)
    env = make_env(jax, ns, layout, name, random_reset=random_reset, max_steps=max_steps, task_id=task_id)
    vreset = jax.jit(jax.vmap(env.reset, in_axes=(0,)))
    vstep = jax.jit(jax.vmap(env.step, in_axes=(0, 0, 0)))
    T = int(np.asarray(actions).shape[0])
    obs, st = vreset(jnp.asarray(np.asarray(reset_keys, dtype=np.uint32)))
    rec = {"r_obs0": np.asarray(obs["agent_0"]), "r_obs1": np.asarray(obs["agent_1"])}
    for f, dt in DYN.items():
        rec["r_" + f] = np.asarray(getattr(st, f)).astype(dt)
    rows = {k: [] for k in ["s_obs0", "s_obs1", "s_reward", "s_shaped0", "s_shaped1", "s_done"] + ["s_" + f for f in DYN]}
    for t in range(T):
        a = np.asarray(actions[t], dtype=np.int32)
        act = {"agent_0": jnp.asarray(a[:, 0]), "agent_1": jnp.asarray(a[:, 1])}
        obs, st, rew, done, info = vstep(jnp.asarray(np.asarray(step_keys[t], dtype=np.uint32)), st, act)
        rows["s_obs0"].append(np.asarray(obs["agent_0"]))
        rows["s_obs1"].append(np.asarray(obs["agent_1"]))
        rows["s_reward"].append(np.asarray(rew["agent_0"], dtype=np.float32))
        rows["s_shaped0"].append(np.asarray(info["shaped_reward"]["agent_0"], dtype=np.float32))
        rows["s_shaped1"].append(np.asarray(info["shaped_reward"]["agent_1"], dtype=np.float32))
        rows["s_done"].append(np.asarray(done["__all__"], dtype=np.bool_))
        for f, dt in DYN.items():
            rows["s_" + f].append(np.asarray(getattr(st, f)).astype(dt))
    rec.update({k: np.stack(v) for k, v in rows.items()})
    return rec


def time_random_rollout(jax, ns, layout, name, *, n_envs, n_steps, random_reset=False, max_steps=400, repeats=3, seed=0):
    """env-steps/s of the reference's jitted, vmapped step in a `lax.scan` with uniform random actions -- the structure of
    baselines/unused/random_actions.py:197-225 (observations ride in the scan carry, as `last_obs` does there), except that
    every env draws its own actions.  Returns (best env-steps/s, seconds of the best run, compile seconds)."""
    jnp, lax, jr = jax.numpy, jax.lax, jax.random
    env = make_env(jax, ns, layout, name, random_reset=random_reset, max_steps=max_steps)
    vstep = jax.vmap(env.step, in_axes=(0, 0, 0))

    def run(rng):
        rng, env_rng = jr.split(rng)
        obsv, env_state = jax.vmap(env.reset, in_axes=(0,))(jr.split(env_rng, n_envs))

        def _env_step(carry, _):
            env_state, last_obs, rng = carry
            rng, key_a, key_s = jr.split(rng, 3)
            a = jr.randint(key_a, (2, n_envs), 0, 6)
            obsv, env_state, reward, done, info = vstep(jr.split(key_s, n_envs), env_state,
                                                        {"agent_0": a[0], "agent_1": a[1]})
            return (env_state, obsv, rng), (reward["agent_0"].sum(), done["__all__"].sum())

        (env_state, obsv, rng), metrics = lax.scan(_env_step, (env_state, obsv, rng), None, length=n_steps)
        return obsv, metrics

    frun = jax.jit(run)
    t0 = time.perf_counter()
    jax.block_until_ready(frun(jr.PRNGKey(seed)))
    compile_s = time.perf_counter() - t0
    best = None
    for r in range(repeats):
        t0 = time.perf_counter()
        jax.block_until_ready(frun(jr.PRNGKey(seed + 1 + r)))
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return n_envs * n_steps / best, best, compile_s
