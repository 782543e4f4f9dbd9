"""JAX side of the XLA FFI binding (csrc/xla_ffi_shim.cc): lets the REFERENCE's `Overcooked` keep its
`reset / step / get_obs` signatures under `jax.jit / jax.vmap / lax.scan` while the work runs in liboc_b200.so.

STATUS: EXPERIMENTAL, NEVER RUN.  Importable only where JAX (>= 0.4.31, `jax.ffi`) is installed -- NOT in this image, so
nothing here is exercised by the test-suite beyond a compile check of the C++ side against a stand-in header; it is the
binding a maintainer of the reference would add.

    import jax_marl
    from jaxovercooked_b200 import jax_ffi
    env = jax_marl.make("overcooked", layout=layout)
    jax_ffi.bind(env)                      # env.reset / env.step / env.get_obs now call the sm_100a kernels
    obs, state = jax.vmap(env.reset)(keys)                                  # one batched custom call
    obs, state, rew, done, info = jax.vmap(env.step, in_axes=(0, 0, 0))(keys, state, actions)
"""
import ctypes as C
import os
import types

import numpy as np

from . import _lib

_FFI_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboc_b200_ffi.so")
_STATE_FIELDS = _lib.STATE_FIELDS          # order == oc_state == the reference's State dataclass
_registered = False


def _register():
    global _registered
    import jax
    if _registered:
        return
    if not os.path.exists(_FFI_LIB):
        raise _lib.OcError(f"{_FFI_LIB} is missing: build it with __graft_entry__.build() on a machine that has JAX")
    C.CDLL(_lib.LIB_PATH, mode=C.RTLD_GLOBAL)              # the shim links against liboc_b200.so
    shim = C.CDLL(_FFI_LIB)
    for name, sym in (("oc_step", "OcStep"), ("oc_step_reset_state", "OcStepResetState"), ("oc_rollout", "OcRollout"),
                      ("oc_reset", "OcReset"), ("oc_get_obs", "OcGetObs"), ("oc_policy_act_cells", "OcPolicyActCells")):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(shim, sym)), platform="CUDA")
    _registered = True


def _layout_handle(env, prng_mode):
    """oc_layout_create for the reference env object (same fields the host mirror passes)."""
    lib = _lib.load()
    lay = env.layout
    keep = {f: np.ascontiguousarray(np.asarray(lay[f], dtype=np.int32).reshape(-1))
            for f in ("wall_idx", "agent_idx", "goal_idx", "pot_idx", "onion_pile_idx", "plate_pile_idx")}
    d = _lib.LayoutDesc()
    d.height, d.width = int(env.height), int(env.width)
    d.n_wall, d.n_goal, d.n_pot = keep["wall_idx"].size, keep["goal_idx"].size, keep["pot_idx"].size
    d.n_onion_pile, d.n_plate_pile = keep["onion_pile_idx"].size, keep["plate_pile_idx"].size
    for f, a in keep.items():
        setattr(d, f, a.ctypes.data_as(C.POINTER(C.c_int32)))
    d.random_reset, d.max_steps, d.task_id, d.prng_mode = int(bool(env.random_reset)), int(env.max_steps), int(env.task_id), prng_mode
    h = C.c_void_p()
    _lib.check(lib.oc_layout_create(C.byref(d), C.byref(h)), "oc_layout_create")
    return h.value, keep


def bind(env, donate_state=True):
    """Replace `env.reset / env.step / env.get_obs` of a reference `Overcooked` instance by FFI custom calls.
    The `State` pytree, dict keys, dtypes and shapes are unchanged, so LogWrapper and every caller keep working."""
    import jax
    import jax.numpy as jnp
    _register()
    part = int(bool(jax.config.jax_threefry_partitionable))
    handle, keep = _layout_handle(env, part)
    env._oc_layout_handle, env._oc_keepalive = handle, keep
    h, w = int(env.height), int(env.width)
    State = None

    def _state_types(lead, G, P):
        shp = {"agent_pos": (2, 2), "agent_dir": (2, 2), "agent_dir_idx": (2,), "agent_inv": (2,), "goal_pos": (G, 2),
               "pot_pos": (P, 2), "wall_map": (h, w), "maze_map": (h + 8, w + 8, 3), "time": (), "terminal": (), "task_id": ()}
        dt = {"agent_pos": jnp.uint32, "agent_dir": jnp.int8, "agent_dir_idx": jnp.int32, "agent_inv": jnp.int32,
              "goal_pos": jnp.uint32, "pot_pos": jnp.uint32, "wall_map": jnp.bool_, "maze_map": jnp.uint8,
              "time": jnp.int32, "terminal": jnp.bool_, "task_id": jnp.int32}
        return [jax.ShapeDtypeStruct(lead + shp[f], dt[f]) for f in _STATE_FIELDS]

    G, P = keep["goal_idx"].size, keep["pot_idx"].size
    obs_t = lambda lead: [jax.ShapeDtypeStruct(lead + (h, w, 26), jnp.uint8)] * 2

    def reset(self, key):
        nonlocal State
        if State is None:
            from jax_marl.environments.overcooked_environment.overcooked import State as _S
            State = _S
        lead = key.shape[:-1]
        res = jax.ffi.ffi_call("oc_reset", _state_types(lead, G, P) + obs_t(lead), vmap_method="broadcast_all")(
            key, layout=np.int64(handle))
        st = State(**dict(zip(_STATE_FIELDS, res[:11])))
        return {"agent_0": res[11], "agent_1": res[12]}, st

    def get_obs(self, state):
        lead = state.time.shape
        res = jax.ffi.ffi_call("oc_get_obs", obs_t(lead), vmap_method="broadcast_all")(
            *[jnp.asarray(getattr(state, f)) for f in _STATE_FIELDS], layout=np.int64(handle))
        return {"agent_0": res[0], "agent_1": res[1]}

    def step(self, key, state, actions, reset_state=None):
        lead = key.shape[:-1]
        outs = (_state_types(lead, G, P) + obs_t(lead) + [jax.ShapeDtypeStruct(lead, jnp.float32)] * 3 +
                [jax.ShapeDtypeStruct(lead, jnp.bool_)])
        aliases = {1 + i: i for i in range(11)} if donate_state else {}
        operands = [key] + [jnp.asarray(getattr(state, f)) for f in _STATE_FIELDS] + [
            jnp.asarray(actions["agent_0"], jnp.int32), jnp.asarray(actions["agent_1"], jnp.int32)]
        target = "oc_step"
        if reset_state is not None:      # multi_agent_env.py:47,55-59: eleven more operands, same results
            operands += [jnp.asarray(getattr(reset_state, f)) for f in _STATE_FIELDS]
            target = "oc_step_reset_state"
        res = jax.ffi.ffi_call(target, outs, vmap_method="broadcast_all", input_output_aliases=aliases)(
            *operands, layout=np.int64(handle))
        st = type(state)(**dict(zip(_STATE_FIELDS, res[:11])))
        obs0, obs1, reward, sh0, sh1, done = res[11:]
        return ({"agent_0": obs0, "agent_1": obs1}, st, {"agent_0": reward, "agent_1": reward},
                {"agent_0": done, "agent_1": done, "__all__": done}, {"shaped_reward": {"agent_0": sh0, "agent_1": sh1}})

    def rollout(self, keys, state, actions):
        """T steps in one custom call (`oc_rollout`): keys u32[T, N, 2], actions {agent: i32[T, N]}; returns what
        `jax.lax.scan` over `jax.vmap(env.step)` would stack: obs u8[T, N, h, w, 26] per agent, reward / shaped f32[T, N],
        done bool[T, N], and the final State (always aliased onto `state`: the kernel steps in place)."""
        T, n = keys.shape[0], keys.shape[1]
        lead = (n,)
        outs = (_state_types(lead, G, P) + [jax.ShapeDtypeStruct((T, n, h, w, 26), jnp.uint8)] * 2 +
                [jax.ShapeDtypeStruct((T, n), jnp.float32)] * 3 + [jax.ShapeDtypeStruct((T, n), jnp.bool_)])
        res = jax.ffi.ffi_call("oc_rollout", outs, input_output_aliases={1 + i: i for i in range(11)})(
            keys, *[jnp.asarray(getattr(state, f)) for f in _STATE_FIELDS],
            jnp.asarray(actions["agent_0"], jnp.int32), jnp.asarray(actions["agent_1"], jnp.int32), layout=np.int64(handle))
        st = type(state)(**dict(zip(_STATE_FIELDS, res[:11])))
        obs0, obs1, reward, sh0, sh1, done = res[11:]
        return ({"agent_0": obs0, "agent_1": obs1}, st, {"agent_0": reward, "agent_1": reward},
                {"agent_0": done, "agent_1": done, "__all__": done}, {"shaped_reward": {"agent_0": sh0, "agent_1": sh1}})

    env.reset = types.MethodType(reset, env)
    env.step = types.MethodType(step, env)
    env.get_obs = types.MethodType(get_obs, env)
    env.rollout = types.MethodType(rollout, env)
    return env


def policy_act_cells(policy_handle, cells, key, action_dim):
    """`action, log_prob, value, logits` for both agents of the N envs whose compact observations are `cells`
    (uint8 [N, cell_stride], as written through `oc_step_extras.cells`): one `oc_policy_forward_cells` custom call.
    `policy_handle` is the integer value of an `oc_policy*` (see include/oc_b200.h, oc_policy_create with a layout)."""
    import jax
    import jax.numpy as jnp
    _register()
    m = 2 * cells.shape[0]
    outs = [jax.ShapeDtypeStruct((m, action_dim), jnp.float32), jax.ShapeDtypeStruct((m,), jnp.float32),
            jax.ShapeDtypeStruct((m,), jnp.int32), jax.ShapeDtypeStruct((m,), jnp.float32)]
    logits, value, action, log_prob = jax.ffi.ffi_call("oc_policy_act_cells", outs)(cells, key, policy=np.int64(policy_handle))
    return action, log_prob, value, logits
