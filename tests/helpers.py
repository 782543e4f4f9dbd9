"""Shared helpers for the parity tests: golden-fixture loading and field-by-field comparison."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DYN_FIELDS = ("agent_pos", "agent_dir", "agent_dir_idx", "agent_inv", "maze_map", "time", "terminal")
STATIC_FIELDS = ("goal_pos", "pot_pos", "wall_map", "task_id")
LOG_FIELDS = ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths")


def golden_cases():
    """Environment trajectories (P_* are the actor-critic fixtures, see policy_cases)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith("P_")]


def policy_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "P_*.npz")))


def load_golden(case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(g["meta"]))
    lay = {k: (np.asarray(v, np.int32) if isinstance(v, list) else v) for k, v in meta["layout"].items()}
    return g, meta, lay


def assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype != b.dtype:
        raise AssertionError(f"{what}: dtype {a.dtype} vs {b.dtype}")
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {bad[0].tolist()}: "
                             f"{a[tuple(bad[0])]} vs {b[tuple(bad[0])]}")


def crafted_state(g):
    """Full batched State (numpy dict) of a crafted fixture: dynamic fields from r_*, static ones broadcast."""
    E = g["r_time"].shape[0]
    st = {f: np.ascontiguousarray(g["r_" + f]) for f in DYN_FIELDS}
    for f in STATIC_FIELDS:
        st[f] = np.ascontiguousarray(np.broadcast_to(g[f], (E,) + g[f].shape))
    return st
