"""Layout descriptors for the Overcooked path (pure host-side Python; no device code here).

Mirrors the reference's layout surface:
  * `overcooked_layouts`  name -> layout dict with the eight fields the env reads
    (reference: jax_marl/environments/overcooked_environment/layouts.py:651-719; the data is extracted
    from the reference by tests/golden/make_layouts.py into layouts_data.json, index order preserved)
  * `layout_grid_to_dict` ASCII grid -> layout dict (reference: layouts.py:79-127)
  * `pad_layouts`         the callers' `pad_observation_space` (reference: baselines/IPPO_CL.py:182-278),
    which re-indexes a task sequence onto a common (max_h, max_w) grid by adding outer walls.

A layout dict holds `height`, `width` (int) and `wall_idx`, `agent_idx`, `goal_idx`, `plate_pile_idx`,
`onion_pile_idx`, `pot_idx` (1-D int32 numpy arrays of flat cell indices, idx = row * width + col).
"""
import json
import os

import numpy as np

LAYOUT_FIELDS = ("wall_idx", "agent_idx", "goal_idx", "plate_pile_idx", "onion_pile_idx", "pot_idx")

_SYMBOL_FIELD = {"W": "wall_idx", "A": "agent_idx", "X": "goal_idx",
                 "B": "plate_pile_idx", "O": "onion_pile_idx", "P": "pot_idx"}


class Layout(dict):
    """dict with identity hash, standing in for flax's FrozenDict (callers use it as a static jit arg)."""

    def __hash__(self):
        return id(self)


def _mk(height, width, **idx):
    d = Layout(height=int(height), width=int(width))
    for f in LAYOUT_FIELDS:
        d[f] = np.asarray(idx.get(f, ()), dtype=np.int32).reshape(-1)
    return d


def layout_grid_to_dict(grid):
    """W wall, A agent, X goal, B plate pile, O onion pile, P pot, space = floor; piles/pots/goals are walls too.

    Follows layouts.py:79-127 including its quirk that `width` is taken from the first row only
    (a longer row aliases into the next row's indices)."""
    rows = grid.split("\n")
    if len(rows[0]) == 0:
        rows = rows[1:]
    if len(rows[-1]) == 0:
        rows = rows[:-1]
    width = len(rows[0])
    acc = {f: [] for f in LAYOUT_FIELDS}
    for i, row in enumerate(rows):
        for j, ch in enumerate(row):
            idx = width * i + j
            if ch in _SYMBOL_FIELD:
                acc[_SYMBOL_FIELD[ch]].append(idx)
            if ch in "XBOP":
                acc["wall_idx"].append(idx)
    return _mk(len(rows), width, **acc)


def _load_registry():
    path = os.path.join(os.path.dirname(__file__), "layouts_data.json")
    with open(path) as fh:
        raw = json.load(fh)
    reg = {}
    for name, d in raw["layouts"].items():
        reg[name] = _mk(d["height"], d["width"], **{f: d[f] for f in LAYOUT_FIELDS})
    groups = {g: {n: reg[n] for n in names} for g, names in raw["groups"].items()}
    return reg, groups


overcooked_layouts, _groups = _load_registry()
hard_layouts = _groups["hard_layouts"]
medium_layouts = _groups["medium_layouts"]
easy_layouts = _groups["easy_layouts"]
padded_layouts = _groups["padded_layouts"]


def pad_layouts(layouts):
    """Pad every layout of a task sequence to the sequence's (max height, max width) with outer walls.

    left = dw // 2, right = dw - left, top = dh // 2, bottom = dh - top; every index moves to
    (row + top) * max_w + col + left; all added cells become walls (appended after the existing walls,
    top rows, bottom rows, then left/right columns row by row) -- IPPO_CL.py:207-270."""
    layouts = list(layouts)
    max_w = max(int(l["width"]) for l in layouts)
    max_h = max(int(l["height"]) for l in layouts)
    out = []
    for lay in layouts:
        w, h = int(lay["width"]), int(lay["height"])
        dw, dh = max_w - w, max_h - h
        left, top = dw // 2, dh // 2
        right, bottom = dw - left, dh - top

        def adj(a):
            a = np.asarray(a, dtype=np.int64)
            return ((a // w + top) * (w + left + right) + (a % w + left)).astype(np.int32)

        new = {f: adj(lay[f]) for f in LAYOUT_FIELDS}
        walls = list(new["wall_idx"])
        for y in range(top):
            walls += [y * max_w + x for x in range(max_w)]
        for y in range(max_h - bottom, max_h):
            walls += [y * max_w + x for x in range(max_w)]
        for y in range(top, max_h - bottom):
            walls += [y * max_w + x for x in range(left)]
            walls += [y * max_w + x for x in range(max_w - right, max_w)]
        new["wall_idx"] = np.asarray(walls, dtype=np.int32)
        out.append(_mk(max_h, max_w, **new))
    return out


CL_SEQUENCE = ("cramped_room", "asymm_advantages", "coord_ring", "forced_coord", "counter_circuit")


def cl_sequence_padded():
    """BASELINE.json config 3: the 5-layout continual-learning sequence padded to a common 5x9 grid."""
    return dict(zip(CL_SEQUENCE, pad_layouts([overcooked_layouts[n] for n in CL_SEQUENCE])))
