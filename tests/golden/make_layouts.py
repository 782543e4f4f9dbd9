"""Extract the reference's layout registry into `jaxovercooked_b200/layouts_data.json`.

Run in the build container (needs /root/reference):  python -m tests.golden.make_layouts
Executes the reference's `layouts.py` (through tests/golden/jaxshim.py) and dumps, for every entry of
`overcooked_layouts` (layouts.py:651-719), the eight fields the environment reads: height, width,
wall_idx, agent_idx, goal_idx, plate_pile_idx, onion_pile_idx, pot_idx -- index ORDER preserved, since
`agent_idx[0]` is agent_0 and list order decides `pot_pos` / `goal_pos` row order in `State`.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from tests.golden import jaxshim  # noqa: E402

FIELDS = ("wall_idx", "agent_idx", "goal_idx", "plate_pile_idx", "onion_pile_idx", "pot_idx")


def main():
    ns = jaxshim.load_reference()
    lm = ns.layouts_module
    out = {"layouts": {}, "groups": {}}
    for name, lay in ns.layouts.items():
        d = {"height": int(lay["height"]), "width": int(lay["width"])}
        for f in FIELDS:
            d[f] = [int(v) for v in np.asarray(lay[f]).reshape(-1)]
        out["layouts"][name] = d
    for g in ("hard_layouts", "medium_layouts", "easy_layouts", "padded_layouts"):
        out["groups"][g] = list(getattr(lm, g).keys())
    path = os.path.join(os.path.dirname(__file__), "..", "..", "jaxovercooked_b200", "layouts_data.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=None, separators=(",", ":"), sort_keys=False)
        fh.write("\n")
    print(f"wrote {len(out['layouts'])} layouts -> {os.path.normpath(path)}")


if __name__ == "__main__":
    main()
