"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/oc_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
The product package (jaxovercooked_b200) never does.

`build()` compiles oracle/oc_oracle.c with gcc (+OpenMP) into oracle/_build/liboc_oracle.so.
State arrays are NumPy, struct-of-arrays with a leading N, same dtypes/shapes as the reference's `State`
(overcooked.py:44-56) and as include/oc_b200.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oc_oracle.c")
_OUT = os.path.join(_HERE, "_build", "liboc_oracle.so")

MAX_CELLS = 256
MAX_SPECIAL = 16
PAD = 4


class CLayout(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32),
                ("n_wall", C.c_int32), ("n_goal", C.c_int32), ("n_pot", C.c_int32),
                ("n_onion_pile", C.c_int32), ("n_plate_pile", C.c_int32),
                ("wall_idx", C.c_int32 * (MAX_CELLS * 2)),
                ("agent_idx", C.c_int32 * 2),
                ("goal_idx", C.c_int32 * MAX_SPECIAL),
                ("pot_idx", C.c_int32 * MAX_SPECIAL),
                ("onion_pile_idx", C.c_int32 * MAX_SPECIAL),
                ("plate_pile_idx", C.c_int32 * MAX_SPECIAL),
                ("random_reset", C.c_int32), ("max_steps", C.c_int32), ("task_id", C.c_int32),
                ("partitionable", C.c_int32)]


class CState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("agent_pos", "agent_dir", "agent_dir_idx", "agent_inv", "goal_pos",
                                          "pot_pos", "wall_map", "maze_map", "time", "terminal", "task_id")]


STATE_FIELDS = tuple(n for n, _ in CState._fields_)


def build(force=False):
    """Compile the C oracle if missing or stale. Returns the .so path."""
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c11", "-Wall", "-o", _OUT, _SRC]
        subprocess.run(cmd, check=True)
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oco_validate.restype = C.c_int
        L.oco_max_threads.restype = C.c_int
        _lib = L
    return _lib


def make_layout(layout, random_reset=False, max_steps=400, task_id=0, partitionable=False):
    cl = CLayout()
    cl.height, cl.width = int(layout["height"]), int(layout["width"])

    def fill(dst, src, cap):
        src = [int(v) for v in np.asarray(src).reshape(-1)]
        if len(src) > cap:
            raise ValueError("layout list too long for the oracle")
        for i, v in enumerate(src):
            dst[i] = v
        return len(src)

    cl.n_wall = fill(cl.wall_idx, layout["wall_idx"], MAX_CELLS * 2)
    if fill(cl.agent_idx, layout["agent_idx"], 2) != 2:
        raise ValueError("agent_idx must hold exactly two cells")
    cl.n_goal = fill(cl.goal_idx, layout["goal_idx"], MAX_SPECIAL)
    cl.n_pot = fill(cl.pot_idx, layout["pot_idx"], MAX_SPECIAL)
    cl.n_onion_pile = fill(cl.onion_pile_idx, layout["onion_pile_idx"], MAX_SPECIAL)
    cl.n_plate_pile = fill(cl.plate_pile_idx, layout["plate_pile_idx"], MAX_SPECIAL)
    cl.random_reset, cl.max_steps = int(bool(random_reset)), int(max_steps)
    cl.task_id, cl.partitionable = int(task_id), int(bool(partitionable))
    rc = lib().oco_validate(C.byref(cl))
    if rc != 0:
        raise ValueError(f"oracle: unsupported layout (code {rc})")
    return cl


def state_shapes(cl, N):
    h, w, G, P = cl.height, cl.width, cl.n_goal, cl.n_pot
    return {"agent_pos": ((N, 2, 2), np.uint32), "agent_dir": ((N, 2, 2), np.int8),
            "agent_dir_idx": ((N, 2), np.int32), "agent_inv": ((N, 2), np.int32),
            "goal_pos": ((N, G, 2), np.uint32), "pot_pos": ((N, P, 2), np.uint32),
            "wall_map": ((N, h, w), np.bool_), "maze_map": ((N, h + 2 * PAD, w + 2 * PAD, 3), np.uint8),
            "time": ((N,), np.int32), "terminal": ((N,), np.bool_), "task_id": ((N,), np.int32)}


def alloc_state(cl, N):
    return {k: np.zeros(shp, dt) for k, (shp, dt) in state_shapes(cl, N).items()}


def _cstate(st):
    cs = CState()
    for f in STATE_FIELDS:
        a = st[f]
        assert a.flags["C_CONTIGUOUS"], f
        setattr(cs, f, a.ctypes.data)
    return cs


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Batched CPU environment with the reference's semantics; mutates `state` dicts in place on step."""

    def __init__(self, layout, random_reset=False, max_steps=400, task_id=0, partitionable=False):
        self.cl = make_layout(layout, random_reset, max_steps, task_id, partitionable)
        self.h, self.w = self.cl.height, self.cl.width
        self.partitionable = bool(partitionable)

    def reset(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 2)
        N = keys.shape[0]
        st = alloc_state(self.cl, N)
        obs0 = np.zeros((N, self.h, self.w, 26), np.uint8)
        obs1 = np.zeros_like(obs0)
        lib().oco_reset(C.byref(self.cl), C.c_int64(N), _p(keys), _cstate(st), _p(obs0), _p(obs1))
        return (obs0, obs1), st

    def get_obs(self, st):
        N = st["time"].shape[0]
        obs0 = np.zeros((N, self.h, self.w, 26), np.uint8)
        obs1 = np.zeros_like(obs0)
        lib().oco_get_obs(C.byref(self.cl), C.c_int64(N), _cstate(st), _p(obs0), _p(obs1))
        return obs0, obs1

    def step(self, keys, st, act0, act1, log=None, out=None):
        """In-place step. `log` = dict of four float32[N,2] arrays (LogWrapper) or None.
        Returns (obs0, obs1), reward, shaped0, shaped1, done."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 2)
        N = keys.shape[0]
        act0 = np.ascontiguousarray(act0, dtype=np.int32)
        act1 = np.ascontiguousarray(act1, dtype=np.int32)
        if out is None:
            out = self.alloc_out(N)
        lw = [None] * 4 if log is None else [log[k] for k in ("episode_returns", "episode_lengths",
                                                              "returned_episode_returns", "returned_episode_lengths")]
        lib().oco_step(C.byref(self.cl), C.c_int64(N), _p(keys), _cstate(st), _p(act0), _p(act1),
                       _p(out["obs0"]), _p(out["obs1"]), _p(out["reward"]), _p(out["shaped0"]), _p(out["shaped1"]),
                       _p(out["done"]), *[_p(a) for a in lw])
        return (out["obs0"], out["obs1"]), out["reward"], out["shaped0"], out["shaped1"], out["done"]

    def alloc_out(self, N):
        return {"obs0": np.zeros((N, self.h, self.w, 26), np.uint8), "obs1": np.zeros((N, self.h, self.w, 26), np.uint8),
                "reward": np.zeros(N, np.float32), "shaped0": np.zeros(N, np.float32),
                "shaped1": np.zeros(N, np.float32), "done": np.zeros(N, np.bool_)}

    @staticmethod
    def new_log(N):
        return {k: np.zeros((N, 2), np.float32) for k in ("episode_returns", "episode_lengths",
                                                           "returned_episode_returns", "returned_episode_lengths")}


def split(key, num, partitionable=False):
    key = np.ascontiguousarray(key, dtype=np.uint32).reshape(2)
    out = np.zeros((num, 2), np.uint32)
    lib().oco_split(_p(key), C.c_int(num), C.c_int(int(partitionable)), _p(out))
    return out


def threefry2x32(k0, k1, x0, x1):
    a, b = C.c_uint32(), C.c_uint32()
    lib().oco_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(x0), C.c_uint32(x1), C.byref(a), C.byref(b))
    return a.value, b.value


def max_threads():
    return int(lib().oco_max_threads())


def set_threads(n):
    lib().oco_set_threads(C.c_int(int(n)))
