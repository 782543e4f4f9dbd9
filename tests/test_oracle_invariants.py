"""SURVEY.md App. A.8 invariants on oracle rollouts for every registered layout (CPU)."""
import numpy as np
import pytest

from oracle import oracle as orc
from jaxovercooked_b200.layouts import overcooked_layouts


@pytest.mark.parametrize("name", list(overcooked_layouts))
def test_invariants_hold_along_a_rollout(name):
    lay = overcooked_layouts[name]
    h, w = int(lay["height"]), int(lay["width"])
    N, T, max_steps = 64, 60, 23
    o = orc.Oracle(lay, random_reset=True, max_steps=max_steps)
    rng = np.random.default_rng(hash(name) % 2**32)
    (obs0, obs1), st = o.reset(rng.integers(0, 2**32, (N, 2), dtype=np.uint32))
    ring = np.ones((h + 8, w + 8), bool)
    ring[4:-4, 4:-4] = False
    walls = np.zeros(h * w, bool)
    walls[np.asarray(lay["wall_idx"])] = True
    for t in range(T):
        a = rng.choice(6, size=(N, 2), p=[0.15, 0.15, 0.15, 0.15, 0.05, 0.35]).astype(np.int32)
        (obs0, obs1), rew, s0, s1, done = o.step(rng.integers(0, 2**32, (N, 2), dtype=np.uint32), st, a[:, 0], a[:, 1])
        mm = st["maze_map"]
        assert (mm[:, ring] == np.array([2, 5, 0], np.uint8)).all()                       # the ring never changes
        assert ((mm[..., 0] == 10).sum(axis=(1, 2)) == 2).all()                           # exactly two agent cells
        pos = st["agent_pos"].astype(int)
        for i in range(2):
            cell = mm[np.arange(N), 4 + pos[:, i, 1], 4 + pos[:, i, 0]]
            assert (cell[:, 0] == 10).all() and (cell[:, 1] == 2 * i).all()
            assert (cell[:, 2] == st["agent_dir_idx"][:, i]).all()
            assert not walls[pos[:, i, 1] * w + pos[:, i, 0]].any()                       # never on a wall
        assert np.isin(st["agent_inv"], [1, 3, 5, 9]).all()
        pots = mm[..., 0] == 8
        assert (mm[..., 2][pots] <= 23).all() and (mm[..., 1][pots] == 7).all()
        assert ((st["time"] >= 0) & (st["time"] < max_steps)).all() and not st["terminal"].any()
        assert np.isin(rew, [0, 20, 40]).all() and np.isin(s0, [0, 3, 5]).all() and np.isin(s1, [0, 3, 5]).all()
        assert obs0.max() <= 19 and (obs0[..., [13, 17, 19, 24]] == 0).all()
        assert (obs0[..., 0].sum(axis=(1, 2)) == 1).all() and (obs0[..., 1].sum(axis=(1, 2)) == 1).all()
        urg = (max_steps - st["time"]) < 40
        assert (obs0[..., 25] == urg[:, None, None]).all()
        assert (obs1[..., 0] == obs0[..., 1]).all() and (obs1[..., 2:6] == obs0[..., 6:10]).all()
        assert (obs1[..., 10:] == obs0[..., 10:]).all()
