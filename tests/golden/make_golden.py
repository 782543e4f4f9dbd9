"""Generate golden trajectories by EXECUTING THE REFERENCE'S OWN SOURCE for the hot path.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python -m tests.golden.make_golden

The reference's `overcooked.py`, `common.py`, `layouts.py`, `multi_agent_env.py` and the `LogWrapper`
from `wrappers/baselines.py` are imported UNMODIFIED through `tests/golden/jaxshim.py` (NumPy stand-ins
for jax/flax/chex, because JAX cannot be installed here) and driven exactly as the callers do
(baselines/IPPO_CL.py:477-478, 535-538): `reset(key)` then repeated `step(key, state, actions)`.
Randomness comes from `oracle/prng.py` (our restatement of jax.random -- not pinned against real JAX).

Each fixture `tests/golden/<case>.npz` holds, for E independent envs and T steps:
    meta (json string): layout name / dict, random_reset, max_steps, task_id, prng mode
    reset_keys u32[E,2], step_keys u32[T,E,2], actions i32[T,E,2]
    r_*  : the reset output        (obs0, obs1, and every dynamic State field), leading [E]
    s_*  : the per-step outputs    (same + reward, shaped0, shaped1, done),     leading [T,E]
    static State fields (goal_pos, pot_pos, wall_map, task_id) once -- asserted constant over time
    lw_* : LogWrapper outputs for the cases that wrap the env (leading [T,E])
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from tests.golden import jaxshim  # noqa: E402
from oracle import prng  # noqa: E402
from jaxovercooked_b200.layouts import overcooked_layouts, cl_sequence_padded, LAYOUT_FIELDS  # noqa: E402

OUT = os.path.dirname(__file__)

# dtype each field has under real JAX with x64 disabled
DYN = {"agent_pos": np.uint32, "agent_dir": np.int8, "agent_dir_idx": np.int32, "agent_inv": np.int32,
       "maze_map": np.uint8, "time": np.int32, "terminal": np.bool_}
STATIC = {"goal_pos": np.uint32, "pot_pos": np.uint32, "wall_map": np.bool_, "task_id": np.int32}

ACTION_P = np.array([0.15, 0.15, 0.15, 0.15, 0.05, 0.35])  # interact-heavy: random play rarely cooks otherwise


def to_ref_layout(ns, lay):
    d = {"height": int(lay["height"]), "width": int(lay["width"])}
    for f in LAYOUT_FIELDS:
        d[f] = jaxshim._array(np.asarray(lay[f], dtype=np.int32))
    return ns.FrozenDict(d)


def rollout(ns, lay, name, *, random_reset, max_steps, task_id, seed, E, T, actions=None, log_wrapper=False,
            partitionable=False):
    jaxshim._Cfg.partitionable = partitionable
    env = ns.Overcooked(layout=to_ref_layout(ns, lay), layout_name=name, random_reset=random_reset,
                        max_steps=max_steps, task_id=task_id)
    wenv = ns.LogWrapper(env) if log_wrapper else env
    rng = prng.PRNGKey(seed)
    arng = np.random.default_rng(seed)
    if actions is None:
        actions = arng.choice(6, size=(T, E, 2), p=ACTION_P).astype(np.int32)
    actions = np.asarray(actions, dtype=np.int32)
    assert actions.shape == (T, E, 2)

    rng, sub = prng.split(rng, 2, partitionable)
    reset_keys = prng.split(sub, E, partitionable)
    step_keys = np.zeros((T, E, 2), np.uint32)
    for t in range(T):
        rng, sub = prng.split(rng, 2, partitionable)
        step_keys[t] = prng.split(sub, E, partitionable)

    rec = {"r_obs0": [], "r_obs1": [], "s_obs0": [], "s_obs1": [], "s_reward": [], "s_shaped0": [],
           "s_shaped1": [], "s_done": []}
    for f in DYN:
        rec["r_" + f] = []
        rec["s_" + f] = []
    lw = {k: [] for k in ("lw_episode_returns", "lw_episode_lengths", "lw_returned_episode_returns",
                          "lw_returned_episode_lengths", "lw_returned_episode")}
    static = None

    def env_state(s):
        return s.env_state if log_wrapper else s

    states = []
    for e in range(E):
        obs, st = wenv.reset(reset_keys[e].view(jaxshim.Arr))
        states.append(st)
        es = env_state(st)
        rec["r_obs0"].append(np.asarray(obs["agent_0"]))
        rec["r_obs1"].append(np.asarray(obs["agent_1"]))
        for f, dt in DYN.items():
            rec["r_" + f].append(np.asarray(getattr(es, f)).astype(dt))
        cur = {f: np.asarray(getattr(es, f)).astype(dt) for f, dt in STATIC.items()}
        if static is None:
            static = cur
        for f in STATIC:
            assert np.array_equal(static[f], cur[f])

    for t in range(T):
        row = {k: [] for k in rec if k.startswith("s_")}
        lrow = {k: [] for k in lw}
        for e in range(E):
            act = {"agent_0": np.int32(actions[t, e, 0]), "agent_1": np.int32(actions[t, e, 1])}
            obs, st, rew, done, info = wenv.step(step_keys[t, e].view(jaxshim.Arr), states[e], act)
            states[e] = st
            es = env_state(st)
            assert float(rew["agent_0"]) == float(rew["agent_1"])
            assert bool(done["agent_0"]) == bool(done["agent_1"]) == bool(done["__all__"])
            row["s_obs0"].append(np.asarray(obs["agent_0"]))
            row["s_obs1"].append(np.asarray(obs["agent_1"]))
            row["s_reward"].append(np.float32(rew["agent_0"]))
            row["s_shaped0"].append(np.float32(info["shaped_reward"]["agent_0"]))
            row["s_shaped1"].append(np.float32(info["shaped_reward"]["agent_1"]))
            row["s_done"].append(bool(done["__all__"]))
            for f, dt in DYN.items():
                row["s_" + f].append(np.asarray(getattr(es, f)).astype(dt))
            for f, dt in STATIC.items():
                assert np.array_equal(static[f], np.asarray(getattr(es, f)).astype(dt)), f
            if log_wrapper:
                lrow["lw_episode_returns"].append(np.asarray(st.episode_returns, np.float32))
                lrow["lw_episode_lengths"].append(np.asarray(st.episode_lengths, np.float32))
                lrow["lw_returned_episode_returns"].append(np.asarray(info["returned_episode_returns"], np.float32))
                lrow["lw_returned_episode_lengths"].append(np.asarray(info["returned_episode_lengths"], np.float32))
                lrow["lw_returned_episode"].append(np.asarray(info["returned_episode"], np.bool_))
        for k, v in row.items():
            rec[k].append(np.stack(v))
        if log_wrapper:
            for k, v in lrow.items():
                lw[k].append(np.stack(v))

    out = {k: np.stack(v) for k, v in rec.items()}
    assert out["s_obs0"].dtype == np.uint8 and out["r_obs0"].dtype == np.uint8
    out.update(static)
    if log_wrapper:
        out.update({k: np.stack(v) for k, v in lw.items()})
    out["reset_keys"] = reset_keys
    out["step_keys"] = step_keys
    out["actions"] = actions
    meta = {"layout_name": name, "random_reset": bool(random_reset), "max_steps": int(max_steps),
            "task_id": int(task_id), "partitionable": bool(partitionable), "seed": int(seed),
            "layout": {"height": int(lay["height"]), "width": int(lay["width"]),
                       **{f: [int(v) for v in lay[f]] for f in LAYOUT_FIELDS}},
            "log_wrapper": bool(log_wrapper)}
    out["meta"] = np.array(json.dumps(meta))
    return out


# hand-derived cramped_room scenario (agent_1 always `stay`); 39 scripted steps then `stay`
SCENARIO_A0 = [3, 5, 2, 0, 5, 3, 5, 2, 0, 5, 3, 5, 2, 0, 5, 1, 3, 1, 5, 0, 2, 0] + [4] * 12 + [5, 1, 2, 1, 5]


def main():
    t0 = time.time()
    ns = jaxshim.load_reference()
    np.seterr(over="ignore")
    cases = {}
    names = list(overcooked_layouts)

    # (A) every registered layout, fixed start (random_reset=False), short episodes so auto-reset fires twice
    for i, n in enumerate(names):
        cases[f"A_{n}"] = dict(lay=overcooked_layouts[n], name=n, random_reset=False, max_steps=19,
                               task_id=i % 3, seed=100 + i, E=2, T=44)
    # (B) every registered layout, random_reset=True (random cells, pot states, inventories)
    for i, n in enumerate(names):
        cases[f"B_{n}"] = dict(lay=overcooked_layouts[n], name=n, random_reset=True, max_steps=23,
                               task_id=0, seed=300 + i, E=2, T=50)
    # (C) cramped_room, default 400-step episode through the LogWrapper (urgency plane, returned_* bookkeeping)
    cases["C_cramped_room_400"] = dict(lay=overcooked_layouts["cramped_room"], name="cramped_room",
                                       random_reset=False, max_steps=400, task_id=0, seed=30, E=2, T=405,
                                       log_wrapper=True)
    # (D) the hand-derived scenario of SURVEY.md App. D, executed by the reference
    T = 402
    a = np.full((T, 1, 2), 4, np.int32)
    a[:len(SCENARIO_A0), 0, 0] = SCENARIO_A0
    cases["D_scenario"] = dict(lay=overcooked_layouts["cramped_room"], name="cramped_room", random_reset=False,
                               max_steps=400, task_id=0, seed=7, E=1, T=T, actions=a, log_wrapper=True)
    # (E) the five continual-learning layouts padded to 5x9 (BASELINE.json config 3)
    for i, (n, lay) in enumerate(cl_sequence_padded().items()):
        cases[f"E_padded_{n}"] = dict(lay=lay, name=n, random_reset=(i % 2 == 1), max_steps=31, task_id=i,
                                      seed=500 + i, E=2, T=70, log_wrapper=True)
    # (F) partitionable-threefry mode (jax >= 0.5 default)
    for i, n in enumerate(("cramped_room", "asymm_advantages", "easy_layout_padded")):
        cases[f"F_part_{n}"] = dict(lay=overcooked_layouts[n], name=n, random_reset=True, max_steps=13,
                                    task_id=0, seed=700 + i, E=2, T=30, partitionable=True)

    # (G) longer interact-heavy play on the five CL layouts (native size): counter drops/pickups, shared cells
    for i, n in enumerate(("cramped_room", "asymm_advantages", "coord_ring", "forced_coord", "counter_circuit")):
        cases[f"G_long_{n}"] = dict(lay=overcooked_layouts[n], name=n, random_reset=True, max_steps=37,
                                    task_id=0, seed=900 + i, E=3, T=200)

    total = 0
    for cname, kw in cases.items():
        out = rollout(ns, **kw)
        path = os.path.join(OUT, cname + ".npz")
        np.savez_compressed(path, **out)
        total += os.path.getsize(path)
    print(f"{len(cases)} cases, {total / 1e6:.2f} MB, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
