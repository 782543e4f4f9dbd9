"""Live differential test: the reference's OWN source (executed through tests/golden/jaxshim.py) against the C
oracle on fresh seeds.  Runs only where /root/reference is mounted (the build container); the committed golden
files are the same comparison frozen for the GPU box."""
import os

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "jax_marl")), reason="reference tree not mounted")


@pytest.mark.parametrize("name,random_reset,seed", [("cramped_room", True, 11), ("asymm_advantages", False, 12),
                                                    ("forced_coord", True, 13), ("easy_layout_padded", True, 14)])
def test_reference_source_vs_oracle_fresh_seed(name, random_reset, seed):
    from tests.golden import jaxshim
    from tests.golden.make_golden import rollout
    from oracle import oracle as orc
    from jaxovercooked_b200.layouts import overcooked_layouts
    from tests.helpers import DYN_FIELDS, assert_same
    ns = jaxshim.load_reference(REF)
    old = np.seterr(over="ignore")
    try:
        g = rollout(ns, overcooked_layouts[name], name, random_reset=random_reset, max_steps=9, task_id=2, seed=seed,
                    E=2, T=24)
    finally:
        np.seterr(**old)
    o = orc.Oracle(overcooked_layouts[name], random_reset=random_reset, max_steps=9, task_id=2)
    (obs0, obs1), st = o.reset(g["reset_keys"])
    assert_same(obs0, g["r_obs0"], "reset obs0")
    for t in range(24):
        (obs0, obs1), rew, sh0, sh1, done = o.step(g["step_keys"][t], st, g["actions"][t, :, 0], g["actions"][t, :, 1])
        assert_same(obs0, g["s_obs0"][t], f"t={t} obs0")
        assert_same(obs1, g["s_obs1"][t], f"t={t} obs1")
        assert_same(rew, g["s_reward"][t], f"t={t} reward")
        assert_same(done, g["s_done"][t], f"t={t} done")
        for f in DYN_FIELDS:
            assert_same(st[f], g["s_" + f][t], f"t={t} {f}")
