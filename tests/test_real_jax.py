"""The reference itself (its unmodified source under jax.jit(jax.vmap(...))) against the CPU oracle, on identical keys and
actions, driven by baseline/ref_jax.py -- the same harness `bench.py --impl reference` uses when JAX is available.

  * test_real_jax_*            need a real `jax` and the reference tree (baseline/_ref or /root/reference): they close the
                               one parity gap the build image leaves open -- jax.random.randint / choice / split of a REAL
                               JAX install, in both `jax_threefry_partitionable` settings -- and skip
                               cleanly where JAX is not installed (this image, the GPU boxes).
  * test_harness_wiring_on_the_shim   runs the SAME harness code with tests/golden/jaxshim.py registered as `jax`
                               (NumPy stand-in; its jax.random is oracle/prng.py), so the harness is known to work before a
                               JAX install ever sees it.  Needs only the reference tree.
"""
import os

import numpy as np
import pytest

from baseline import ref_jax
from tests.helpers import DYN_FIELDS, assert_same

needs_ref = pytest.mark.skipif(ref_jax.find_reference() is None, reason="reference tree not mounted")

CASES = [("cramped_room", False, 21), ("cramped_room", True, 22), ("coord_ring", True, 23),
         ("bottleneck_large", True, 24)]


def _inputs(E, T, seed, partitionable):
    from oracle import prng
    rng = prng.PRNGKey(seed)
    g = np.random.default_rng(seed)
    rng, sub = prng.split(rng, 2, partitionable)
    reset_keys = prng.split(sub, E, partitionable)
    step_keys = np.zeros((T, E, 2), np.uint32)
    for t in range(T):
        rng, sub = prng.split(rng, 2, partitionable)
        step_keys[t] = prng.split(sub, E, partitionable)
    actions = g.choice(6, size=(T, E, 2), p=[0.15, 0.15, 0.15, 0.15, 0.05, 0.35]).astype(np.int32)
    return reset_keys, step_keys, actions


def _diff_against_oracle(jax, ns, name, random_reset, seed, partitionable, E=3, T=30, max_steps=11):
    from oracle import oracle as orc
    from jaxovercooked_b200.layouts import overcooked_layouts
    lay = overcooked_layouts[name]
    reset_keys, step_keys, actions = _inputs(E, T, seed, partitionable)
    old = np.seterr(over="ignore")
    try:
        g = ref_jax.rollout(jax, ns, lay, name, random_reset=random_reset, max_steps=max_steps, reset_keys=reset_keys,
                            step_keys=step_keys, actions=actions, task_id=1, partitionable=partitionable)
    finally:
        np.seterr(**old)
    o = orc.Oracle(lay, random_reset=random_reset, max_steps=max_steps, task_id=1, partitionable=partitionable)
    (obs0, obs1), st = o.reset(reset_keys)
    assert_same(obs0, g["r_obs0"], "reset obs0")
    assert_same(obs1, g["r_obs1"], "reset obs1")
    for f in DYN_FIELDS:
        assert_same(st[f], g["r_" + f], f"reset {f}")
    for t in range(T):
        (obs0, obs1), rew, sh0, sh1, done = o.step(step_keys[t], st, actions[t, :, 0], actions[t, :, 1])
        assert_same(obs0, g["s_obs0"][t], f"t={t} obs0")
        assert_same(obs1, g["s_obs1"][t], f"t={t} obs1")
        assert_same(rew, g["s_reward"][t], f"t={t} reward")
        assert_same(sh0, g["s_shaped0"][t], f"t={t} shaped0")
        assert_same(sh1, g["s_shaped1"][t], f"t={t} shaped1")
        assert_same(done, g["s_done"][t], f"t={t} done")
        for f in DYN_FIELDS:
            assert_same(st[f], g["s_" + f][t], f"t={t} {f}")


@needs_ref
@pytest.mark.parametrize("name,random_reset,seed", CASES)
def test_harness_wiring_on_the_shim(name, random_reset, seed):
    if ref_jax.import_jax() is not None:
        pytest.skip("a real jax is importable: test_real_jax_* run the harness on it instead")
    from tests.golden import jaxshim
    import sys
    jaxshim.load_reference(ref_jax.find_reference())      # registers the stand-in `jax` and the reference modules
    ns = ref_jax.load_reference()
    _diff_against_oracle(sys.modules["jax"], ns, name, random_reset, seed, partitionable=False, E=2, T=24)


@needs_ref
def test_harness_timing_leg_on_the_shim():
    """bench.py's reference arm on the stand-in: the scan-of-vmapped-steps program runs and returns a rate."""
    if ref_jax.import_jax() is not None:
        pytest.skip("a real jax is importable")
    from tests.golden import jaxshim
    import sys
    from jaxovercooked_b200.layouts import overcooked_layouts
    jaxshim.load_reference(ref_jax.find_reference())
    ns = ref_jax.load_reference()
    old = np.seterr(over="ignore")
    try:
        rate, best, compile_s = ref_jax.time_random_rollout(sys.modules["jax"], ns, overcooked_layouts["cramped_room"],
                                                            "cramped_room", n_envs=2, n_steps=3, repeats=1)
    finally:
        np.seterr(**old)
    assert rate > 0 and best > 0


@needs_ref
@pytest.mark.parametrize("partitionable", [False, True])
@pytest.mark.parametrize("name,random_reset,seed", CASES)
def test_real_jax_env_vs_oracle(name, random_reset, seed, partitionable):
    jax = ref_jax.import_jax()
    if jax is None:
        pytest.skip("jax is not installed (it cannot be, in this image): parity of jax.random draws stays unpinned")
    ns = ref_jax.load_reference()
    _diff_against_oracle(jax, ns, name, random_reset, seed, partitionable)


@pytest.mark.parametrize("partitionable", [False, True])
def test_real_jax_random_known_answers(partitionable):
    """jax.random.split / randint / choice(p, replace=False) / categorical of a real JAX against oracle/prng.py."""
    jax = ref_jax.import_jax()
    if jax is None:
        pytest.skip("jax is not installed")
    from oracle import prng
    jnp, jr = jax.numpy, jax.random
    ref_jax.set_threefry_partitionable(jax, partitionable)
    for seed in (0, 1, 30, 12345):
        key = np.asarray(jr.PRNGKey(seed), dtype=np.uint32)
        assert_same(key, prng.PRNGKey(seed), "PRNGKey")
        for num in (2, 3, 16, 17):
            assert_same(np.asarray(jr.split(jnp.asarray(key), num), dtype=np.uint32), prng.split(key, num, partitionable),
                        f"split {num}")
        for n, span in ((2, 4), (1, 24), (3, 24), (7, 24)):
            got = np.asarray(jr.randint(jnp.asarray(key), (n,), 0, span))
            assert_same(got.astype(np.int32), prng.randint(key, n, 0, span, partitionable).astype(np.int32), f"randint {n} {span}")
        for C in (20, 25, 99):
            mask = (np.random.default_rng(seed + C).random(C) < 0.4).astype(np.float32)
            mask[:2] = 1
            got = np.asarray(jr.choice(jnp.asarray(key), jnp.arange(C), (2,), replace=False, p=jnp.asarray(mask)))
            want = prng.choice_no_replace_p(key, C, 2, mask, partitionable)
            assert_same(got.astype(np.int64), np.asarray(want).astype(np.int64), f"choice {C}")
        # distrax.Categorical(logits).sample(seed=key) == jax.random.categorical(key, logits) (architectures/mlp.py:57)
        from oracle import policy_oracle as po
        logits = np.random.default_rng(seed).normal(size=(9, 6)).astype(np.float32)
        got = np.asarray(jr.categorical(jnp.asarray(key), jnp.asarray(logits)))
        assert_same(got.astype(np.int32), np.asarray(po.sample(key, logits, partitionable)[0]).astype(np.int32), "categorical")
