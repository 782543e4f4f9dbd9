"""Host-side mirror of the reference's Overcooked environment interface, backed by the sm_100a kernels.

Same names, argument meaning and outputs as the reference (jax_marl/environments/overcooked_environment/
overcooked.py:69-104 constructor, :136 reset, :106 / multi_agent_env.py:41-69 step, :250 get_obs; State :44-56),
with two differences that follow from leaving JAX:

  * arrays are torch CUDA tensors (torch is only the allocator / stream provider here), and
  * batching is native instead of `jax.vmap`: every method takes any leading batch shape -- `key[..., 2]`,
    `actions[agent][...]`, State fields with the same leading dims -- and `()` is the reference's un-vmapped
    call.  `jaxovercooked_b200.vmap(env.step)` is provided so call sites written as
    `jax.vmap(env.step, in_axes=(0, 0, 0))(keys, state, actions)` (baselines/IPPO_CL.py:538) keep their shape.

All compute goes through the C ABI in include/oc_b200.h; there is no CPU or PyTorch fallback.
"""
import ctypes as C
import dataclasses
from enum import IntEnum
from typing import Any

import numpy as np
import torch

from . import _lib
from .layouts import Layout, overcooked_layouts, LAYOUT_FIELDS

PAD = 4
N_CHANNELS = 26

# reference constants (overcooked.py:23-31, 59-66; common.py:10-22)
BASE_REW_SHAPING_PARAMS = {"PLACEMENT_IN_POT_REW": 3, "PLATE_PICKUP_REWARD": 3, "SOUP_PICKUP_REWARD": 5,
                           "DROP_COUNTER_REWARD": 0, "DISH_DISP_DISTANCE_REW": 0, "POT_DISTANCE_REW": 0,
                           "SOUP_DISTANCE_REW": 0}
POT_EMPTY_STATUS, POT_FULL_STATUS, POT_READY_STATUS, MAX_ONIONS_IN_POT = 23, 20, 0, 3
URGENCY_CUTOFF, DELIVERY_REWARD = 40, 20
OBJECT_TO_INDEX = {"unseen": 0, "empty": 1, "wall": 2, "onion": 3, "onion_pile": 4, "plate": 5, "plate_pile": 6,
                   "goal": 7, "pot": 8, "dish": 9, "agent": 10}


class Actions(IntEnum):
    up = 0
    down = 1
    right = 2
    left = 3
    stay = 4
    interact = 5
    done = 6


@dataclasses.dataclass
class State:
    """overcooked.py:44-56.  Per-env dtypes/shapes: agent_pos u32[2,2] (x,y), agent_dir i8[2,2],
    agent_dir_idx i32[2], agent_inv i32[2], goal_pos u32[G,2], pot_pos u32[P,2], wall_map bool[h,w],
    maze_map u8[h+8,w+8,3], time i32[], terminal bool[], task_id i32[]; any leading batch dims."""
    agent_pos: Any
    agent_dir: Any
    agent_dir_idx: Any
    agent_inv: Any
    goal_pos: Any
    pot_pos: Any
    wall_map: Any
    maze_map: Any
    time: Any
    terminal: Any
    task_id: Any

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    def clone(self):
        return State(**{f.name: getattr(self, f.name).clone() for f in dataclasses.fields(self)})


_STATE_DTYPES = {"agent_pos": torch.uint32, "agent_dir": torch.int8, "agent_dir_idx": torch.int32,
                 "agent_inv": torch.int32, "goal_pos": torch.uint32, "pot_pos": torch.uint32,
                 "wall_map": torch.bool, "maze_map": torch.uint8, "time": torch.int32, "terminal": torch.bool,
                 "task_id": torch.int32}


class Discrete:
    """jax_marl/environments/spaces.py:19-44 (metadata only)."""

    def __init__(self, num_categories, dtype=torch.int32):
        self.n = num_categories
        self.shape = ()
        self.dtype = dtype


class Box:
    """jax_marl/environments/spaces.py:71-... (metadata only)."""

    def __init__(self, low, high, shape, dtype=torch.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


def _ptr(t):
    return None if t is None else t.data_ptr()


class Overcooked:
    """Vanilla Overcooked (overcooked.py:69), batched natively on one CUDA device."""

    def __init__(self, layout=None, layout_name="cramped_room", random_reset=False, max_steps=400, task_id=0,
                 *, device=None, prng_impl="threefry_legacy"):
        if layout is None:   # the reference dereferences layout before its own fallback (overcooked.py:84,89)
            layout = overcooked_layouts["cramped_room"]
        if not torch.cuda.is_available():
            raise _lib.OcError("jaxovercooked_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.height = int(layout["height"])
        self.width = int(layout["width"])
        self.obs_shape = (self.width, self.height, N_CHANNELS)   # sic: (W, H, 26); arrays are (H, W, 26)
        self.agent_view_size = 5
        self.layout = layout
        self.layout_name = layout_name
        self.agents = ["agent_0", "agent_1"]
        self.num_agents = 2
        self.action_set = torch.tensor([a.value for a in list(Actions)[:6]], dtype=torch.int32)
        self.random_reset = bool(random_reset)
        self.max_steps = int(max_steps)
        self.task_id = int(task_id)
        if prng_impl not in ("threefry_legacy", "threefry_partitionable"):
            raise ValueError("prng_impl must be 'threefry_legacy' (jax < 0.5 default) or 'threefry_partitionable'")
        self.prng_mode = _lib.PRNG_LEGACY if prng_impl == "threefry_legacy" else _lib.PRNG_PARTITIONABLE

        lib = _lib.load()
        self._keep = {f: np.ascontiguousarray(np.asarray(layout[f], dtype=np.int32).reshape(-1)) for f in LAYOUT_FIELDS}
        if self._keep["agent_idx"].size != 2:
            raise ValueError("layout['agent_idx'] must hold exactly two cells")
        d = _lib.LayoutDesc()
        d.height, d.width = self.height, self.width
        d.n_wall = self._keep["wall_idx"].size
        d.n_goal = self._keep["goal_idx"].size
        d.n_pot = self._keep["pot_idx"].size
        d.n_onion_pile = self._keep["onion_pile_idx"].size
        d.n_plate_pile = self._keep["plate_pile_idx"].size
        for f in LAYOUT_FIELDS:
            setattr(d, f, self._keep[f].ctypes.data_as(C.POINTER(C.c_int32)))
        d.random_reset, d.max_steps, d.task_id, d.prng_mode = int(self.random_reset), self.max_steps, self.task_id, self.prng_mode
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.oc_layout_create(C.byref(d), C.byref(handle)), "oc_layout_create")
        self._handle = handle
        self._lib = lib
        info = _lib.LayoutInfo()
        _lib.check(lib.oc_layout_query(handle, C.byref(info)), "oc_layout_query")
        self.info = info
        self.n_goal, self.n_pot = int(info.n_goal), int(info.n_pot)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.oc_layout_destroy(h)
            except Exception:
                pass
            self._handle = None

    # ------------------------------------------------------------------ shapes / buffers
    def state_shapes(self, batch=()):
        h, w, G, P = self.height, self.width, self.n_goal, self.n_pot
        b = tuple(batch)
        return {"agent_pos": b + (2, 2), "agent_dir": b + (2, 2), "agent_dir_idx": b + (2,), "agent_inv": b + (2,),
                "goal_pos": b + (G, 2), "pot_pos": b + (P, 2), "wall_map": b + (h, w),
                "maze_map": b + (h + 2 * PAD, w + 2 * PAD, 3), "time": b, "terminal": b, "task_id": b}

    def empty_state(self, batch=()):
        return State(**{k: torch.empty(s, dtype=_STATE_DTYPES[k], device=self.device)
                        for k, s in self.state_shapes(batch).items()})

    def _state_ptrs(self, st, batch):
        shapes = self.state_shapes(batch)
        sp = _lib.StatePtrs()
        for f in _lib.STATE_FIELDS:
            t = getattr(st, f)
            if t.dtype != _STATE_DTYPES[f] or tuple(t.shape) != shapes[f] or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"State.{f}: expected contiguous {_STATE_DTYPES[f]}{list(shapes[f])} on {self.device}, "
                                 f"got {t.dtype}{list(t.shape)} on {t.device}")
            setattr(sp, f, t.data_ptr())
        return sp

    def _check(self, t, what, dtypes, shape):
        """The kernels write through raw pointers: a wrong dtype, shape, device or a non-contiguous view would corrupt
        memory silently, so every caller-provided buffer is checked before its address crosses the C ABI."""
        dtypes = dtypes if isinstance(dtypes, tuple) else (dtypes,)
        if (not torch.is_tensor(t) or t.dtype not in dtypes or tuple(t.shape) != tuple(shape) or not t.is_contiguous()
                or t.device != self.device):
            got = f"{t.dtype}{list(t.shape)} on {t.device}, contiguous={t.is_contiguous()}" if torch.is_tensor(t) else type(t).__name__
            raise ValueError(f"{what}: expected a contiguous {' or '.join(str(d) for d in dtypes)}{list(shape)} tensor on "
                             f"{self.device}, got {got}")
        return t

    def _check_out(self, out, batch):
        b = tuple(batch)
        img = b + (self.height, self.width, N_CHANNELS)
        if out.get("obs_all") is not None:     # both views of an env side by side (oc_step_extras.obs_interleaved)
            self._check(out["obs_all"], "out['obs_all']", torch.uint8, b + (2, self.height, self.width, N_CHANNELS))
            if out.get("cells") is not None:
                raise ValueError("out['obs_all'] cannot be combined with out['cells']")
        else:
            self._check(out["obs0"], "out['obs0']", torch.uint8, img)
            self._check(out["obs1"], "out['obs1']", torch.uint8, img)
        for k in ("reward", "shaped0", "shaped1"):
            self._check(out[k], f"out['{k}']", torch.float32, b)
        self._check(out["done"], "out['done']", (torch.bool, torch.uint8), b)
        if out.get("cells") is not None:
            self._check(out["cells"], "out['cells']", torch.uint8, b + (self.info.cell_stride,))

    def _check_log(self, log, stats, batch):
        b2 = tuple(batch) + (self.num_agents,)
        if log is not None:
            for f in ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths"):
                if log.get(f) is None:
                    raise ValueError(f"log['{f}'] is missing: the four LogWrapper trackers go together")
                self._check(log[f], f"log['{f}']", torch.float32, b2)
            if log.get("returned_episode") is not None:
                self._check(log["returned_episode"], "log['returned_episode']", (torch.bool, torch.uint8), b2)
        if stats is not None:
            self._check(stats, "stats", torch.int64, (8,))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _key(self, key):
        if key.dtype != torch.uint32 or key.shape[-1] != 2 or key.device != self.device:
            raise ValueError("key must be a uint32 CUDA tensor of shape [..., 2] (raw threefry key data)")
        return key.contiguous()

    def _obs_buffers(self, batch):
        shp = tuple(batch) + (self.height, self.width, N_CHANNELS)
        return (torch.empty(shp, dtype=torch.uint8, device=self.device),
                torch.empty(shp, dtype=torch.uint8, device=self.device))

    # ------------------------------------------------------------------ the reference's three calls
    def reset(self, key, *, cells=None):
        """obs, state = env.reset(key)   (overcooked.py:136-248).  key: u32[..., 2].
        `cells` (uint8 [N, cell_stride], `empty_cells`) also receives the compact form of the first observations."""
        key = self._key(key)
        batch = tuple(key.shape[:-1])
        n = int(np.prod(batch)) if batch else 1
        if cells is not None:
            self._check(cells, "cells", torch.uint8, batch + (self.info.cell_stride,))
        st = self.empty_state(batch)
        obs0, obs1 = self._obs_buffers(batch)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.oc_reset(self._handle, n, key.data_ptr(), self._state_ptrs(st, batch),
                                          obs0.data_ptr(), obs1.data_ptr(), self._stream()), "oc_reset")
            if cells is not None:
                _lib.check(self._lib.oc_get_obs_cells(self._handle, n, self._state_ptrs(st, batch), obs0.data_ptr(),
                                                      obs1.data_ptr(), cells.data_ptr(), self._stream()), "oc_get_obs_cells")
        return {"agent_0": obs0, "agent_1": obs1}, st

    def step(self, key, state, actions, reset_state=None, *, donate=False, out=None, log=None, stats=None):
        """obs, state, rewards, dones, infos = env.step(key, state, actions[, reset_state])
        (multi_agent_env.py:41-69: split key, step_env, auto-reset on done).

        donate=True updates `state`'s tensors in place and returns the same State (what XLA does for a donated
        scan carry); the default allocates a fresh State.  `out` may carry preallocated result tensors
        ({"obs0","obs1","reward","shaped0","shaped1","done"}); `log` / `stats` are used by LogWrapper."""
        lazy = None
        if hasattr(key, "materialize"):        # random.LazySplit: per-env keys derived inside the kernel
            lazy, batch, key_ptr = key, (key.count,), None
            if lazy.key.device != self.device:
                raise ValueError("LazySplit key lives on another device")
        else:
            key = self._key(key)
            batch, key_ptr = tuple(key.shape[:-1]), key.data_ptr()
        n = int(np.prod(batch)) if batch else 1
        a0 = self._action(actions["agent_0"], batch)
        a1 = self._action(actions["agent_1"], batch)
        sp_in = self._state_ptrs(state, batch)
        if donate:
            new_state, sp_out = state, sp_in
        else:
            new_state = self.empty_state(batch)
            sp_out = self._state_ptrs(new_state, batch)
        rs = None
        if reset_state is not None:
            rs = C.byref(self._state_ptrs(reset_state, batch))
        if out is None:
            obs0, obs1 = self._obs_buffers(batch)
            reward = torch.empty(batch, dtype=torch.float32, device=self.device)
            sh0 = torch.empty(batch, dtype=torch.float32, device=self.device)
            sh1 = torch.empty(batch, dtype=torch.float32, device=self.device)
            done = torch.empty(batch, dtype=torch.bool, device=self.device)
        else:
            self._check_out(out, batch)
            reward, sh0, sh1, done = (out[k] for k in ("reward", "shaped0", "shaped1", "done"))
            if out.get("obs_all") is not None:
                obs_all = out["obs_all"]
                nb = len(batch)
                obs0, obs1 = obs_all.select(nb, 0), obs_all.select(nb, 1)     # strided views of the one array
            else:
                obs0, obs1 = out["obs0"], out["obs1"]
        self._check_log(log, stats, batch)
        ex = None
        cells = out.get("cells") if out is not None else None
        obs_all = out.get("obs_all") if out is not None else None
        if log is not None or stats is not None or lazy is not None or cells is not None or obs_all is not None:
            e = _lib.StepExtras()
            if obs_all is not None:
                e.obs_interleaved = 1
            if cells is not None:
                e.cells = cells.data_ptr()
            if lazy is not None:
                e.split_key, e.split_num, e.split_first = lazy.key.data_ptr(), lazy.num, lazy.first
            if log is not None:
                for f in ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths",
                          "returned_episode"):
                    setattr(e, f, _ptr(log.get(f)))
            if stats is not None:
                e.stats = stats.data_ptr()
            ex = C.byref(e)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.oc_step(self._handle, n, key_ptr, sp_in, a0.data_ptr(), a1.data_ptr(), rs,
                                         sp_out, obs0.data_ptr(), None if obs_all is not None else obs1.data_ptr(),
                                         reward.data_ptr(), sh0.data_ptr(),
                                         sh1.data_ptr(), done.data_ptr(), ex, self._stream()), "oc_step")
        obs = {"agent_0": obs0, "agent_1": obs1}
        if obs_all is not None:
            obs["__all__"] = obs_all.view(tuple(batch) + (-1,))               # wrappers/baselines.py:196, no copy
        rewards = {"agent_0": reward, "agent_1": reward}                       # overcooked.py:124
        dones = {"agent_0": done, "agent_1": done, "__all__": done}            # :126
        infos = {"shaped_reward": {"agent_0": sh0, "agent_1": sh1}}            # :125, :133
        return obs, new_state, rewards, dones, infos

    def rollout(self, keys, state, actions, *, out=None, log=None, stats=None):
        """T consecutive `env.step` calls on the same batch in ONE launch (`oc_rollout`): the callers'
        `jax.lax.scan(_env_step, ..., length=T)` over `jax.vmap(env.step)` with the actions given
        (baselines/unused/random_actions.py:197-225).  `state` is stepped IN PLACE (a donated scan carry).

          keys     uint32[T, N, 2] per-step per-env keys, or uint32[T, 2]: step t uses split(keys[t], N) derived in the kernel
                   (or a (keys[T, 2], num, first) tuple for one shard of a global batch of `num` envs)
          actions  {"agent_0": int32[T, N], "agent_1": int32[T, N]}, or a uint8[T, N] tensor (agent_0 | agent_1 << 4)
          out      optional preallocated results: "obs_all" uint8[T, N, 2, h, w, 26] (the VDN global state, both views of an env side
                   by side), or "obs" uint8[T, 2, N, h, w, 26] (the learners' layout: out["obs"][t].view(2N, -1)
                   is batchify(obs) of step t) or "obs0"/"obs1" uint8[T, N, h, w, 26]; "reward"/"shaped0"/"shaped1" float32[T, N],
                   "done" bool[T, N]; or "results8" uint8[T, N] (packed) instead of the four; "cells" uint8[T, N, cell_stride]
          log      LogWrapper trackers [N, 2] (updated T times, in place); log["returned_episode"], if given, is [T, N, 2]
        Returns (obs, state, rewards, dones, infos) with a leading T on every array."""
        dev, h, w = self.device, self.height, self.width
        e = _lib.StepExtras()
        key_ptr = None
        split = None
        if isinstance(keys, tuple):
            keys, num, first = keys
            split = (int(num), int(first))
        if not torch.is_tensor(keys) or keys.dtype != torch.uint32 or keys.device != dev or not keys.is_contiguous():
            raise ValueError("keys must be a contiguous uint32 tensor on the env's device")
        if keys.dim() == 3:
            T, n = int(keys.shape[0]), int(keys.shape[1])
            if keys.shape[2] != 2 or split is not None:
                raise ValueError("keys must be uint32[T, N, 2] or uint32[T, 2]")
            key_ptr = keys.data_ptr()
        elif keys.dim() == 2 and keys.shape[1] == 2:
            T = int(keys.shape[0])
            n = int(state.time.shape[0])
            e.split_key = keys.data_ptr()
            e.split_num, e.split_first = split if split is not None else (n, 0)
        else:
            raise ValueError("keys must be uint32[T, N, 2] or uint32[T, 2]")
        batch = (n,)
        sp = self._state_ptrs(state, batch)
        io = _lib.RolloutIO()
        if torch.is_tensor(actions):
            self._check(actions, "actions", torch.uint8, (T, n))
            io.actions8 = actions.data_ptr()
        else:
            a0 = self._check(actions["agent_0"], "actions['agent_0']", torch.int32, (T, n))
            a1 = self._check(actions["agent_1"], "actions['agent_1']", torch.int32, (T, n))
            io.act0, io.act1 = a0.data_ptr(), a1.data_ptr()
            if n >= 32768:     # large batches: the library packs the actions to one byte per env and step first (oc_rollout_io)
                scratch8 = torch.empty((T, n), dtype=torch.uint8, device=dev)
                io.actions8_scratch = scratch8.data_ptr()
        out = dict(out) if out is not None else {}
        img = (h, w, N_CHANNELS)
        obs_all = out.get("obs_all")
        if obs_all is not None:       # uint8[T, N, 2, h, w, 26]: both views of an env side by side (oc_rollout_io.obs_interleaved)
            self._check(obs_all, "out['obs_all']", torch.uint8, (T, n, 2) + img)
            if out.get("cells") is not None:
                raise ValueError("out['obs_all'] cannot be combined with out['cells']")
            obs0, obs1 = obs_all[:, :, 0], obs_all[:, :, 1]
            stride = 2 * n * h * w * N_CHANNELS
            io.obs_interleaved = 1
        elif out.get("obs0") is not None:
            obs0 = self._check(out["obs0"], "out['obs0']", torch.uint8, (T, n) + img)
            obs1 = self._check(out["obs1"], "out['obs1']", torch.uint8, (T, n) + img)
            if T > 1 and obs0.stride(0) != obs1.stride(0):
                raise ValueError("out['obs0'] and out['obs1'] need the same step stride")
            stride = n * h * w * N_CHANNELS
        else:
            obs = out.get("obs")
            if obs is None:
                obs = torch.empty((T, 2, n) + img, dtype=torch.uint8, device=dev)
            self._check(obs, "out['obs']", torch.uint8, (T, 2, n) + img)
            obs0, obs1 = obs[:, 0], obs[:, 1]
            stride = 2 * n * h * w * N_CHANNELS
        io.obs0, io.obs1, io.obs_step_stride = obs0.data_ptr(), (None if obs_all is not None else obs1.data_ptr()), stride
        res8 = out.get("results8")
        reward = sh0 = sh1 = done = None
        if res8 is not None:
            self._check(res8, "out['results8']", torch.uint8, (T, n))
            io.results8 = res8.data_ptr()
        else:
            reward = out.get("reward") if out.get("reward") is not None else torch.empty((T, n), dtype=torch.float32, device=dev)
            sh0 = out.get("shaped0") if out.get("shaped0") is not None else torch.empty((T, n), dtype=torch.float32, device=dev)
            sh1 = out.get("shaped1") if out.get("shaped1") is not None else torch.empty((T, n), dtype=torch.float32, device=dev)
            done = out.get("done") if out.get("done") is not None else torch.empty((T, n), dtype=torch.bool, device=dev)
            for k, t in (("reward", reward), ("shaped0", sh0), ("shaped1", sh1)):
                self._check(t, f"out['{k}']", torch.float32, (T, n))
            self._check(done, "out['done']", (torch.bool, torch.uint8), (T, n))
            io.reward, io.shaped0, io.shaped1, io.done = reward.data_ptr(), sh0.data_ptr(), sh1.data_ptr(), done.data_ptr()
        if out.get("cells") is not None:
            self._check(out["cells"], "out['cells']", torch.uint8, (T, n, self.info.cell_stride))
            e.cells = out["cells"].data_ptr()
        if log is not None:
            for f in ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths"):
                if log.get(f) is None:
                    raise ValueError(f"log['{f}'] is missing: the four LogWrapper trackers go together")
                setattr(e, f, self._check(log[f], f"log['{f}']", torch.float32, (n, 2)).data_ptr())
            if log.get("returned_episode") is not None:
                e.returned_episode = self._check(log["returned_episode"], "log['returned_episode']",
                                                 (torch.bool, torch.uint8), (T, n, 2)).data_ptr()
        if stats is not None:
            e.stats = self._check(stats, "stats", torch.int64, (8,)).data_ptr()
        with torch.cuda.device(dev):
            _lib.check(self._lib.oc_rollout(self._handle, n, T, key_ptr, sp, C.byref(io), C.byref(e), self._stream()),
                       "oc_rollout")
        obs_d = {"agent_0": obs0, "agent_1": obs1}
        if obs_all is not None:
            obs_d["__all__"] = obs_all.view(T, n, -1)
        if res8 is not None:
            return obs_d, state, {"results8": res8}, {"results8": res8}, {"results8": res8}
        rewards = {"agent_0": reward, "agent_1": reward}
        dones = {"agent_0": done, "agent_1": done, "__all__": done}
        infos = {"shaped_reward": {"agent_0": sh0, "agent_1": sh1}}
        return obs_d, state, rewards, dones, infos

    @staticmethod
    def unpack_results8(res8):
        """Decode the packed results of `rollout(out={"results8": ...})` / HostRollout(packed=True):
        (reward f32, shaped0 f32, shaped1 f32, done bool), same shape as `res8` (torch tensor or numpy array)."""
        lut = res8.new_tensor([0.0, 3.0, 5.0, 0.0], dtype=torch.float32) if torch.is_tensor(res8) else np.array([0, 3, 5, 0], np.float32)
        r = res8.to(torch.int64) if torch.is_tensor(res8) else res8.astype(np.int64)
        reward = (r & 3) * 20
        reward = reward.to(torch.float32) if torch.is_tensor(res8) else reward.astype(np.float32)
        return reward, lut[(r >> 2) & 3], lut[(r >> 4) & 3], ((r >> 6) & 1) != 0

    # ------------------------------------------------------------------ host-policy fast path
    def host_stepper(self, state, n_global=None, first=0, stream=None):
        """Bind one env batch for repeated `oc_step_host` calls: actions arrive in pinned host memory, rewards /
        shaped rewards / dones are written to pinned host memory, observations and state stay on the device.
        Everything is enqueued on `stream` (default: a new stream), so several batches overlap copy and compute."""
        return HostStepper(self, state, n_global, first, stream)

    def get_obs(self, state, cells=None):
        """overcooked.py:250-364.  `cells` (uint8 [N, cell_stride]) also receives the compact form."""
        batch = tuple(state.time.shape)
        n = int(np.prod(batch)) if batch else 1
        if cells is not None:
            self._check(cells, "cells", torch.uint8, batch + (self.info.cell_stride,))
        obs0, obs1 = self._obs_buffers(batch)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.oc_get_obs_cells(self._handle, n, self._state_ptrs(state, batch), obs0.data_ptr(),
                                                  obs1.data_ptr(), cells.data_ptr() if cells is not None else None,
                                                  self._stream()), "oc_get_obs")
        return {"agent_0": obs0, "agent_1": obs1}

    def get_obs_all(self, state, out=None):
        """Both agents' observations as ONE uint8[..., 2, h, w, 26] array (`oc_get_obs_all`): `[..., 0]` / `[..., 1]` are
        get_obs()["agent_0"] / ["agent_1"], and flattened per env it is the CTRolloutManager's obs["__all__"]."""
        batch = tuple(state.time.shape)
        n = int(np.prod(batch)) if batch else 1
        shp = batch + (2, self.height, self.width, N_CHANNELS)
        obs_all = torch.empty(shp, dtype=torch.uint8, device=self.device) if out is None else self._check(out, "out", torch.uint8, shp)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.oc_get_obs_all(self._handle, n, self._state_ptrs(state, batch), obs_all.data_ptr(),
                                                self._stream()), "oc_get_obs_all")
        return obs_all

    def empty_cells(self, n):
        """Buffer for the compact observations of `n` environments (see include/oc_b200.h, oc_step_extras.cells)."""
        return torch.full((n, self.info.cell_stride), 255, dtype=torch.uint8, device=self.device)

    def cell_tables(self):
        """(records uint8[68, 26], scan_cells int32[n_scan]): how to expand the compact form back into observations."""
        rec = np.zeros((68, 26), np.uint8)
        sc = np.zeros((self.info.n_scan,), np.int32)
        _lib.check(self._lib.oc_layout_cell_tables(self._handle, rec.ctypes.data, sc.ctypes.data), "oc_layout_cell_tables")
        return rec, sc

    def _action(self, a, batch):
        if not torch.is_tensor(a):
            a = torch.as_tensor(a, dtype=torch.int32, device=self.device)
        if a.dtype != torch.int32:
            a = a.to(torch.int32)
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        if tuple(a.shape) != tuple(batch):
            raise ValueError(f"action shape {tuple(a.shape)} does not match the key batch {tuple(batch)}")
        return a.contiguous()

    # ------------------------------------------------------------------ metadata the callers read
    def obs_template(self):
        """The layout's static observation (uint8 numpy [h, w, 26]): what `get_obs` returns minus everything that can
        change.  `policy.ActorCritic.bind(..., obs_template=env.obs_template())` evaluates layer 1 relative to it."""
        import numpy as np
        out = np.empty((self.height, self.width, 26), np.uint8)
        _lib.check(self._lib.oc_layout_obs_template(self._handle, out.ctypes.data), "oc_layout_obs_template")
        return out

    def is_terminal(self, state):
        return (state.time >= self.max_steps) | state.terminal                 # overcooked.py:644-647

    @property
    def name(self):
        return self.layout_name

    @property
    def num_actions(self):
        return len(self.action_set)

    def action_space(self, agent_id=""):
        return Discrete(len(self.action_set), dtype=torch.uint32)

    def observation_space(self, agent_id=""):
        return Box(0, 255, self.obs_shape)

    def state_space(self):
        h, w, v = self.height, self.width, self.agent_view_size
        return {"agent_pos": Box(0, max(w, h), (2,), torch.uint32), "agent_dir": Discrete(4),
                "goal_pos": Box(0, max(w, h), (2,), torch.uint32),
                "maze_map": Box(0, 255, (w + v, h + v, 3), torch.uint32),
                "time": Discrete(self.max_steps), "terminal": Discrete(2)}


class HostStepper:
    """One env batch bound to `oc_step_host` (include/oc_b200.h): a single C call per step enqueues the H2D copy of
    the actions, the in-place step, and the D2H copy of the results.  `actions` is the pinned int32[2,N] host
    tensor to fill before `step()`; after `wait()` the pinned views `reward / shaped0 / shaped1 / done` hold the
    step's results.  Per-env keys are derived in the kernel from `split(key, n_global)[first + i]`."""

    def __init__(self, env, state, n_global=None, first=0, stream=None):
        batch = tuple(state.time.shape)
        if len(batch) != 1 or batch[0] % 4:
            raise ValueError("host_stepper needs a 1-D batch whose size is a multiple of 4")
        self.env, self.state, self.n = env, state, batch[0]
        self.n_global = self.n if n_global is None else int(n_global)
        self.first = int(first)
        n, dev = self.n, env.device
        self.stream = torch.cuda.Stream(dev) if stream is None else stream
        self.actions = torch.zeros((2, n), dtype=torch.int32).pin_memory()
        self._results = torch.zeros(13 * n, dtype=torch.uint8).pin_memory()
        self.reward = self._results[:4 * n].view(torch.float32)
        self.shaped0 = self._results[4 * n:8 * n].view(torch.float32)
        self.shaped1 = self._results[8 * n:12 * n].view(torch.float32)
        self.done = self._results[12 * n:].view(torch.bool)
        self._scratch = torch.empty(21 * n + 16, dtype=torch.uint8, device=dev)
        self.obs = {"agent_0": torch.empty((n, env.height, env.width, N_CHANNELS), dtype=torch.uint8, device=dev),
                    "agent_1": torch.empty((n, env.height, env.width, N_CHANNELS), dtype=torch.uint8, device=dev)}
        self._sp = env._state_ptrs(state, batch)
        ex = _lib.StepExtras()
        ex.split_num, ex.split_first = self.n_global, self.first
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        handle = C.c_void_p()
        _lib.check(env._lib.oc_host_stepper_create(env._handle, n, env._state_ptrs(state, batch), self._results.data_ptr(),
                                                   self._scratch.data_ptr(), self.obs["agent_0"].data_ptr(),
                                                   self.obs["agent_1"].data_ptr(), C.byref(ex),
                                                   C.c_void_p(self.stream.cuda_stream), C.byref(handle)),
                   "oc_host_stepper_create")
        self._h = handle
        self._step, self._wait, self._destroy = (env._lib.oc_host_stepper_step, env._lib.oc_host_stepper_wait,
                                                 env._lib.oc_host_stepper_destroy)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._destroy(h)
            except Exception:
                pass
            self._h = None

    def step(self, key, actions=None):
        """Enqueue one step; `key` is the step's subkey (uint32[2] CUDA tensor), split per env inside the kernel.
        `actions`: a pinned int32[2,N] host tensor to use instead of `self.actions` (no host-side copy)."""
        act = self.actions if actions is None else actions
        if actions is not None and (not torch.is_tensor(act) or act.dtype != torch.int32 or tuple(act.shape) != (2, self.n)
                                    or not act.is_contiguous() or act.device.type != "cpu" or not act.is_pinned()):
            raise ValueError(f"actions must be a contiguous pinned int32 host tensor of shape (2, {self.n})")
        if not torch.is_tensor(key) or key.dtype != torch.uint32 or tuple(key.shape) != (2,) or key.device != self.env.device:
            raise ValueError("key must be a uint32[2] tensor on the env's device (the step's subkey)")
        self._keepalive = (key, act)
        rc = self._step(self._h, key.data_ptr(), act.data_ptr())
        if rc:
            _lib.check(rc, "oc_host_stepper_step")

    def wait(self):
        """Block until this batch's last enqueued step has delivered its results to the pinned host views."""
        rc = self._wait(self._h)
        if rc:
            _lib.check(rc, "oc_host_stepper_wait")


class HostRollout:
    """One env batch bound to `oc_rollout_host` (include/oc_b200.h): ONE C call enqueues the host-to-device copy of T steps'
    actions, one T-step launch and the device-to-host copy of the T steps' results on this batch's stream.
      packed=False (reference dtypes): `actions` pinned int32[2, T, N]; results pinned views reward / shaped0 / shaped1
                   float32[T, N], done bool[T, N] -- 21 bytes per env-step over PCIe
      packed=True:  `actions` pinned uint8[T, N] (agent_0 | agent_1 << 4); `results8` pinned uint8[T, N]
                   (Overcooked.unpack_results8) -- 2 bytes per env-step
    Observations stay on the device in `obs` uint8[T, 2, N, h, w, 26] (the learners' layout); step t's per-env keys are
    split(keys[t], n_global)[first + i], derived in the kernel."""

    def __init__(self, env, state, num_steps, n_global=None, first=0, stream=None, packed=False, obs=None):
        batch = tuple(state.time.shape)
        if len(batch) != 1 or (batch[0] * num_steps) % 4:
            raise ValueError("HostRollout needs a 1-D batch with num_steps * N a multiple of 4")
        self.env, self.state, self.n, self.T, self.packed = env, state, batch[0], int(num_steps), bool(packed)
        self.n_global = self.n if n_global is None else int(n_global)
        self.first = int(first)
        n, T, dev = self.n, self.T, env.device
        self.stream = torch.cuda.Stream(dev) if stream is None else stream
        self._actions = None               # pinned host action buffer, allocated on first use (callers may bring their own)
        if packed:
            self.results8 = torch.zeros((T, n), dtype=torch.uint8).pin_memory()
            self._results = self.results8
            self._scratch = torch.empty(2 * T * n + 16, dtype=torch.uint8, device=dev)
        else:
            self._results = torch.zeros(13 * T * n, dtype=torch.uint8).pin_memory()
            tn = T * n
            self.reward = self._results[:4 * tn].view(torch.float32).view(T, n)
            self.shaped0 = self._results[4 * tn:8 * tn].view(torch.float32).view(T, n)
            self.shaped1 = self._results[8 * tn:12 * tn].view(torch.float32).view(T, n)
            self.done = self._results[12 * tn:].view(torch.bool).view(T, n)
            self._scratch = torch.empty(21 * T * n + 16, dtype=torch.uint8, device=dev)
        img = (env.height, env.width, N_CHANNELS)
        self.obs = torch.empty((T, 2, n) + img, dtype=torch.uint8, device=dev) if obs is None else env._check(
            obs, "obs", torch.uint8, (T, 2, n) + img)
        self._sp = env._state_ptrs(state, batch)
        self._ex = _lib.StepExtras()
        self._ex.split_num, self._ex.split_first = self.n_global, self.first
        self._stride = 2 * n * env.height * env.width * N_CHANNELS
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        self._call = env._lib.oc_rollout_host

    @property
    def actions(self):
        """The pinned host tensor to fill with the T steps' actions (uint8[T, N] packed, int32[2, T, N] otherwise)."""
        if self._actions is None:
            shape, dt = ((self.T, self.n), torch.uint8) if self.packed else ((2, self.T, self.n), torch.int32)
            self._actions = torch.zeros(shape, dtype=dt).pin_memory()
        return self._actions

    def run(self, keys, actions=None, log=None, stats=None):
        """Enqueue T steps.  keys: uint32[T, 2] CUDA tensor (the T per-step subkeys); `actions`: a pinned host tensor to use
        instead of `self.actions`."""
        env, T, n = self.env, self.T, self.n
        act = self.actions if actions is None else actions
        want = ((T, n), torch.uint8) if self.packed else ((2, T, n), torch.int32)
        if (not torch.is_tensor(act) or tuple(act.shape) != want[0] or act.dtype != want[1] or not act.is_contiguous()
                or act.device.type != "cpu" or not act.is_pinned()):
            raise ValueError(f"actions must be a contiguous pinned {want[1]} host tensor of shape {want[0]}")
        if (not torch.is_tensor(keys) or keys.dtype != torch.uint32 or tuple(keys.shape) != (T, 2) or keys.device != env.device
                or not keys.is_contiguous()):
            raise ValueError("keys must be a contiguous uint32[T, 2] tensor on the env's device")
        env._check_log(None if log is None else {k: v for k, v in log.items() if k != "returned_episode"}, stats, (n,))
        ex = self._ex
        ex.split_key = keys.data_ptr()
        for f in ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths"):
            setattr(ex, f, log[f].data_ptr() if log is not None else None)      # (never keep a previous call's pointers)
        ex.stats = stats.data_ptr() if stats is not None else None
        self._keepalive = (keys, act, log, stats)
        with torch.cuda.device(env.device):
            rc = self._call(env._handle, n, T, self._sp, _lib.WIRE_PACKED if self.packed else _lib.WIRE_REFERENCE,
                            act.data_ptr(), self._results.data_ptr(), self._scratch.data_ptr(), self.obs[:, 0].data_ptr(),
                            self.obs[:, 1].data_ptr(), self._stride, C.byref(ex), C.c_void_p(self.stream.cuda_stream))
        if rc:
            _lib.check(rc, "oc_rollout_host")

    def wait(self):
        self.stream.synchronize()


registered_envs = ["overcooked"]


def make(env_id, **env_kwargs):
    """jax_marl/registration.py:5-28 for the one environment family in scope."""
    if env_id not in registered_envs:
        raise ValueError(f"{env_id} is not in registered jaxmarl environments.")
    return Overcooked(**env_kwargs)


def vmap(fn, in_axes=0, out_axes=0):
    """Stand-in for `jax.vmap` at the reference's call sites: the env methods are already batched over leading
    dimensions, so mapping over axis 0 of every argument is the function itself."""
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,)
    if any(a != 0 for a in axes) or out_axes != 0:
        raise NotImplementedError("only in_axes=0 / out_axes=0 (the reference's usage) is supported")
    return fn
