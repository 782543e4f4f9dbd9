"""The C oracle against the golden trajectories recorded from the reference's own source
(tests/golden/make_golden.py).  Bit-exact: every State field, both observations, rewards, shaped rewards,
dones and the LogWrapper bookkeeping, at reset and after every step."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import (DYN_FIELDS, STATIC_FIELDS, LOG_FIELDS, assert_same, crafted_state, golden_cases,
                           load_golden)


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_trajectory(case):
    g, meta, lay = load_golden(case)
    o = orc.Oracle(lay, random_reset=meta["random_reset"], max_steps=meta["max_steps"], task_id=meta["task_id"],
                   partitionable=meta["partitionable"])
    if meta.get("crafted"):      # hand-built states (tests/golden/make_interact_matrix.py): get_obs, then one step
        st = crafted_state(g)
        obs0, obs1 = o.get_obs(st)
    else:
        (obs0, obs1), st = o.reset(g["reset_keys"])
    assert_same(obs0, g["r_obs0"], "reset obs0")
    assert_same(obs1, g["r_obs1"], "reset obs1")
    for f in DYN_FIELDS:
        assert_same(st[f], g["r_" + f], f"reset {f}")
    E = g["reset_keys"].shape[0]
    for f in STATIC_FIELDS:
        for e in range(E):
            assert_same(st[f][e], g[f], f"reset {f}[{e}]")
    log = o.new_log(E) if meta["log_wrapper"] else None
    T = g["actions"].shape[0]
    for t in range(T):
        (obs0, obs1), rew, sh0, sh1, done = o.step(g["step_keys"][t], st, g["actions"][t, :, 0],
                                                   g["actions"][t, :, 1], log=log)
        assert_same(obs0, g["s_obs0"][t], f"t={t} obs0")
        assert_same(obs1, g["s_obs1"][t], f"t={t} obs1")
        assert_same(rew, g["s_reward"][t], f"t={t} reward")
        assert_same(sh0, g["s_shaped0"][t], f"t={t} shaped0")
        assert_same(sh1, g["s_shaped1"][t], f"t={t} shaped1")
        assert_same(done, g["s_done"][t], f"t={t} done")
        for f in DYN_FIELDS:
            assert_same(st[f], g["s_" + f][t], f"t={t} {f}")
        for f in STATIC_FIELDS:
            for e in range(E):
                assert_same(st[f][e], g[f], f"t={t} {f}[{e}]")
        if log is not None:
            for f in LOG_FIELDS:
                assert_same(log[f], g["lw_" + f][t], f"t={t} log {f}")
            assert_same(np.stack([done, done], 1), g["lw_returned_episode"][t], f"t={t} returned_episode")
