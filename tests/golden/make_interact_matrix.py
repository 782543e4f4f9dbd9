"""Crafted single-step fixtures: every (faced object x inventory x pot status) combination, executed by the
REFERENCE's own `env.step` (through tests/golden/jaxshim.py).  Complements make_golden.py, whose random rollouts
rarely reach some branches of `process_interact` (overcooked.py:516-642).

    python -m tests.golden.make_interact_matrix      (build container only; needs /root/reference)

Writes tests/golden/X_interact_{cramped_room,coord_ring}.npz in the same format as the trajectory fixtures, with
T = 1 and `r_*` holding the CRAFTED initial state of each case instead of a reset state (meta["crafted"] = true).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from tests.golden import jaxshim  # noqa: E402
from tests.golden.make_golden import DYN, STATIC, to_ref_layout  # noqa: E402
from oracle import prng  # noqa: E402
from jaxovercooked_b200.layouts import overcooked_layouts, LAYOUT_FIELDS  # noqa: E402

OUT = os.path.dirname(__file__)
VEC = {1: (1, 0, 0), 2: (2, 5, 0), 3: (3, 4, 0), 5: (5, 6, 0), 9: (9, 6, 0)}
DIRS = {0: (0, -1), 1: (0, 1), 2: (1, 0), 3: (-1, 0)}


def craft(ns, env, base, pos, dirs, invs, cells, time=0):
    """base: a reset State; returns it with agents moved/turned, inventories set and extra map cells written."""
    A = jaxshim._array
    m = np.array(base.maze_map, copy=True)
    for i in range(2):                                   # erase the reset agents, draw the crafted ones
        x, y = int(base.agent_pos[i][0]), int(base.agent_pos[i][1])
        m[4 + y, 4 + x] = (1, 0, 0)
    for (x, y), v in cells.items():
        m[4 + y, 4 + x] = v
    for i in range(2):
        x, y = pos[i]
        m[4 + y, 4 + x] = (10, 2 * i, dirs[i])
    return base.replace(agent_pos=A(np.array(pos, np.uint32)), agent_dir=A(np.array([DIRS[d] for d in dirs], np.int8)),
                        agent_dir_idx=A(np.array(dirs, np.int32)), agent_inv=A(np.array(invs, np.int32)),
                        maze_map=A(m.astype(np.uint8)), time=time)


def run_cases(ns, name, cases):
    lay = overcooked_layouts[name]
    env = ns.Overcooked(layout=to_ref_layout(ns, lay), layout_name=name, max_steps=400)
    _, base = env.reset(prng.PRNGKey(0).view(jaxshim.Arr))
    E = len(cases)
    rec = {k: [] for k in ["r_obs0", "r_obs1", "s_obs0", "s_obs1", "s_reward", "s_shaped0", "s_shaped1", "s_done"]}
    for f in DYN:
        rec["r_" + f], rec["s_" + f] = [], []
    static = {f: np.asarray(getattr(base, f)).astype(dt) for f, dt in STATIC.items()}
    keys = prng.split(prng.PRNGKey(1), E)
    acts = np.zeros((1, E, 2), np.int32)
    for e, c in enumerate(cases):
        st = craft(ns, env, base, c["pos"], c["dirs"], c["invs"], c.get("cells", {}), c.get("time", 0))
        o = env.get_obs(st)
        rec["r_obs0"].append(np.asarray(o["agent_0"])); rec["r_obs1"].append(np.asarray(o["agent_1"]))
        for f, dt in DYN.items():
            rec["r_" + f].append(np.asarray(getattr(st, f)).astype(dt))
        acts[0, e] = c["acts"]
        obs, st2, rew, done, info = env.step(keys[e].view(jaxshim.Arr), st,
                                             {"agent_0": np.int32(c["acts"][0]), "agent_1": np.int32(c["acts"][1])})
        rec["s_obs0"].append(np.asarray(obs["agent_0"])); rec["s_obs1"].append(np.asarray(obs["agent_1"]))
        rec["s_reward"].append(np.float32(rew["agent_0"]))
        rec["s_shaped0"].append(np.float32(info["shaped_reward"]["agent_0"]))
        rec["s_shaped1"].append(np.float32(info["shaped_reward"]["agent_1"]))
        rec["s_done"].append(bool(done["__all__"]))
        for f, dt in DYN.items():
            rec["s_" + f].append(np.asarray(getattr(st2, f)).astype(dt))
    out = {k: (np.stack(v)[None] if k.startswith("s_") else np.stack(v)) for k, v in rec.items()}
    out.update(static)
    out["reset_keys"] = np.zeros((E, 2), np.uint32)
    out["step_keys"] = keys[None]
    out["actions"] = acts
    meta = {"layout_name": name, "random_reset": False, "max_steps": 400, "task_id": 0, "partitionable": False,
            "seed": 0, "log_wrapper": False, "crafted": True,
            "layout": {"height": int(lay["height"]), "width": int(lay["width"]),
                       **{f: [int(v) for v in lay[f]] for f in LAYOUT_FIELDS}}}
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, f"X_interact_{name}.npz"), **out)
    return E


def main():
    ns = jaxshim.load_reference()
    np.seterr(over="ignore")
    invs_all = (1, 3, 5, 9)
    # ---- cramped_room: agent_0 in front of the pot / piles / goal / a counter / agent_1; agent_1 parked at (3,1)
    cr = []
    for inv in invs_all:
        for status in (0, 1, 7, 19, 20, 21, 22, 23):               # pot at (2,0), faced from (2,1) looking north
            for time in (0, 370):                                   # 370: urgency plane on after the step
                cr.append(dict(pos=[(2, 1), (3, 2)], dirs=[0, 1], invs=[inv, 1], acts=[5, 4],
                               cells={(2, 0): (8, 7, status)}, time=time))
        cr.append(dict(pos=[(1, 1), (3, 1)], dirs=[3, 2], invs=[inv, inv], acts=[5, 5]))          # both at onion piles
        cr.append(dict(pos=[(1, 2), (3, 2)], dirs=[1, 1], invs=[inv, inv], acts=[5, 5]))          # plate pile / goal
        cr.append(dict(pos=[(2, 1), (3, 1)], dirs=[2, 3], invs=[inv, 1], acts=[5, 5]))            # facing each other
        for obj in (2, 3, 5, 9):                                    # counter (0,2) faced from (1,2); (4,2) from (3,2)
            cr.append(dict(pos=[(1, 2), (3, 2)], dirs=[3, 2], invs=[inv, 1], acts=[5, 5],
                           cells={(0, 2): VEC[obj], (4, 2): VEC[obj]}))
            cr.append(dict(pos=[(1, 2), (3, 2)], dirs=[3, 2], invs=[1, inv], acts=[4, 5],
                           cells={(0, 2): VEC[obj], (4, 2): VEC[obj]}))
    # movement corner cases: collisions, swaps, bounces while turning
    for a0 in range(6):
        for a1 in range(6):
            cr.append(dict(pos=[(1, 1), (2, 1)], dirs=[2, 3], invs=[1, 3], acts=[a0, a1]))       # adjacent, facing
            cr.append(dict(pos=[(1, 1), (3, 1)], dirs=[2, 3], invs=[5, 9], acts=[a0, a1]))       # one cell apart
            cr.append(dict(pos=[(1, 2), (2, 1)], dirs=[0, 1], invs=[1, 1], acts=[a0, a1]))       # diagonal
    n1 = run_cases(ns, "cramped_room", cr)
    # ---- coord_ring: both agents face the SAME counter (2,2) from west and east: Alice-before-Bob ordering
    ring = []
    for obj in (2, 3, 5, 9):
        for ia in invs_all:
            for ib in invs_all:
                for acts in ((5, 5), (5, 4), (4, 5)):
                    ring.append(dict(pos=[(1, 2), (3, 2)], dirs=[2, 3], invs=[ia, ib], acts=list(acts),
                                     cells={(2, 2): VEC[obj]}))
                    ring.append(dict(pos=[(3, 2), (1, 2)], dirs=[3, 2], invs=[ia, ib], acts=list(acts),
                                     cells={(2, 2): VEC[obj]}))
    n2 = run_cases(ns, "coord_ring", ring)
    print(f"crafted cases: cramped_room {n1}, coord_ring {n2}")


if __name__ == "__main__":
    main()
