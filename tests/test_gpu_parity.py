"""CUDA path vs oracle / golden, through the C ABI (host mirror -> ctypes -> liboc_b200.so).  Bit-exact for
every State field, both observations, rewards, shaped rewards, dones and LogWrapper bookkeeping."""
import numpy as np
import pytest
import torch

from tests.helpers import (DYN_FIELDS, STATIC_FIELDS, LOG_FIELDS, assert_same, crafted_state, golden_cases,
                           load_golden)

pytestmark = pytest.mark.gpu


def _mods():
    import jaxovercooked_b200 as ocb
    from oracle import oracle as orc
    from tests import gpu_helpers as gh
    return ocb, orc, gh


ACTION_P = np.array([0.15, 0.15, 0.15, 0.15, 0.05, 0.35])


@pytest.mark.parametrize("case", golden_cases())
def test_cuda_matches_reference_golden(case):
    """The CUDA path against trajectories recorded from the reference's own source (no oracle in the loop)."""
    ocb, orc, gh = _mods()
    g, meta, lay = load_golden(case)
    env = ocb.make("overcooked", layout=lay, random_reset=meta["random_reset"], max_steps=meta["max_steps"],
                   task_id=meta["task_id"],
                   prng_impl="threefry_partitionable" if meta["partitionable"] else "threefry_legacy")
    wenv = ocb.LogWrapper(env) if meta["log_wrapper"] else env
    if meta.get("crafted"):      # hand-built states: oc_get_obs on them, then one oc_step
        st = gh.numpy_state_to_dev(env, crafted_state(g))
        obs = env.get_obs(st)
    else:
        obs, st = wenv.reset(gh.keys_to_dev(g["reset_keys"]))
    es = st.env_state if meta["log_wrapper"] else st
    assert_same(obs["agent_0"].cpu().numpy(), g["r_obs0"], "reset obs0")
    assert_same(obs["agent_1"].cpu().numpy(), g["r_obs1"], "reset obs1")
    got = gh.state_to_numpy(es)
    E = g["reset_keys"].shape[0]
    for f in DYN_FIELDS:
        assert_same(got[f], g["r_" + f], f"reset {f}")
    for f in STATIC_FIELDS:
        for e in range(E):
            assert_same(got[f][e], g[f], f"reset {f}[{e}]")
    T = g["actions"].shape[0]
    for t in range(T):
        act = {"agent_0": gh.to_dev(g["actions"][t, :, 0]), "agent_1": gh.to_dev(g["actions"][t, :, 1])}
        obs, st, rew, done, info = wenv.step(gh.keys_to_dev(g["step_keys"][t]), st, act)
        es = st.env_state if meta["log_wrapper"] else st
        assert_same(obs["agent_0"].cpu().numpy(), g["s_obs0"][t], f"t={t} obs0")
        assert_same(obs["agent_1"].cpu().numpy(), g["s_obs1"][t], f"t={t} obs1")
        assert_same(rew["agent_0"].cpu().numpy(), g["s_reward"][t], f"t={t} reward")
        assert_same(info["shaped_reward"]["agent_0"].cpu().numpy(), g["s_shaped0"][t], f"t={t} shaped0")
        assert_same(info["shaped_reward"]["agent_1"].cpu().numpy(), g["s_shaped1"][t], f"t={t} shaped1")
        assert_same(done["__all__"].cpu().numpy(), g["s_done"][t], f"t={t} done")
        got = gh.state_to_numpy(es)
        for f in DYN_FIELDS:
            assert_same(got[f], g["s_" + f][t], f"t={t} {f}")
        for f in STATIC_FIELDS:
            for e in range(E):
                assert_same(got[f][e], g[f], f"t={t} {f}[{e}]")
        if meta["log_wrapper"]:
            for f in LOG_FIELDS:
                assert_same(getattr(st, f).cpu().numpy(), g["lw_" + f][t], f"t={t} log {f}")
            assert_same(info["returned_episode"].cpu().numpy(), g["lw_returned_episode"][t], f"t={t} returned_episode")


def _rollout_compare(lay, *, random_reset, max_steps, N, T, seed, part=False, every=1, donate=True, log=False,
                     time_offsets=False):
    ocb, orc, gh = _mods()
    from oracle import prng
    env = ocb.make("overcooked", layout=lay, random_reset=random_reset, max_steps=max_steps,
                   prng_impl="threefry_partitionable" if part else "threefry_legacy")
    ref = orc.Oracle(lay, random_reset=random_reset, max_steps=max_steps, partitionable=part)
    wenv = ocb.LogWrapper(env) if log else env
    rng = np.random.default_rng(seed)
    rk = rng.integers(0, 2**32, (N, 2), dtype=np.uint32)
    obs, st = wenv.reset(gh.keys_to_dev(rk))
    (r0, r1), rst = ref.reset(rk)
    rlog = ref.new_log(N) if log else None
    assert_same(obs["agent_0"].cpu().numpy(), r0, "reset obs0")
    assert_same(obs["agent_1"].cpu().numpy(), r1, "reset obs1")
    gh.compare_state(st.env_state if log else st, rst, "reset")
    if time_offsets:     # envs at different points of their episodes: urgency and resets hit different slots of a tile
        off = rng.integers(0, max_steps, N).astype(np.int32)
        rst["time"][:] = off
        (st.env_state if log else st).time.copy_(gh.to_dev(off))
    for t in range(T):
        sk = rng.integers(0, 2**32, (N, 2), dtype=np.uint32)
        a = rng.choice(6, size=(N, 2), p=ACTION_P).astype(np.int32)
        act = {"agent_0": gh.to_dev(a[:, 0]), "agent_1": gh.to_dev(a[:, 1])}
        obs, st, rew, done, info = wenv.step(gh.keys_to_dev(sk), st, act, donate=donate)
        (r0, r1), rr, s0, s1, dd = ref.step(sk, rst, a[:, 0], a[:, 1], log=rlog)
        if t % every == 0 or t == T - 1:
            assert_same(obs["agent_0"].cpu().numpy(), r0, f"t={t} obs0")
            assert_same(obs["agent_1"].cpu().numpy(), r1, f"t={t} obs1")
            gh.compare_state(st.env_state if log else st, rst, f"t={t}")
        assert_same(rew["agent_0"].cpu().numpy(), rr, f"t={t} reward")
        assert_same(info["shaped_reward"]["agent_0"].cpu().numpy(), s0, f"t={t} shaped0")
        assert_same(info["shaped_reward"]["agent_1"].cpu().numpy(), s1, f"t={t} shaped1")
        assert_same(done["__all__"].cpu().numpy(), dd, f"t={t} done")
        if log:
            for f in LOG_FIELDS:
                assert_same(getattr(st, f).cpu().numpy(), rlog[f], f"t={t} log {f}")


def _layout_names():
    from jaxovercooked_b200.layouts import overcooked_layouts
    return list(overcooked_layouts)


@pytest.mark.parametrize("random_reset", [False, True])
@pytest.mark.parametrize("name", _layout_names())
def test_all_layouts_vs_oracle(name, random_reset):
    """Every registered layout, ragged batch (not a multiple of the tile), short episodes -> several auto-resets."""
    from jaxovercooked_b200.layouts import overcooked_layouts
    _rollout_compare(overcooked_layouts[name], random_reset=random_reset, max_steps=17, N=203, T=40,
                     seed=hash(name) % 1000 + int(random_reset), every=3)


@pytest.mark.parametrize("max_ctas", ["1", "3"])
@pytest.mark.parametrize("name,random_reset", [("cramped_room", False), ("cramped_room", True), ("coord_ring", True),
                                               ("counter_circuit", False), ("bottleneck_large", True),
                                               ("easy_layout_padded", False)])
def test_many_tiles_per_warp(name, random_reset, max_ctas, monkeypatch):
    """The warp's staging image of the observations persists from tile to tile (only what differs from the static
    image is re-written, and un-written for the next tile).  Small batches give every warp at most one tile, so this
    caps the grid (OC_B200_MAX_CTAS) to push dozens of tiles through each warp, with per-env episode phases mixed so
    that urgency flags and auto-resets differ between the tiles that reuse a staging slot."""
    from jaxovercooked_b200.layouts import overcooked_layouts
    monkeypatch.setenv("OC_B200_MAX_CTAS", max_ctas)
    _rollout_compare(overcooked_layouts[name], random_reset=random_reset, max_steps=61, N=777, T=90,
                     seed=7 + int(max_ctas), every=2, time_offsets=True, log=(max_ctas == "3"))


@pytest.mark.parametrize("part", [False, True])
def test_cl_padded_layouts_logwrapper(part):
    from jaxovercooked_b200.layouts import cl_sequence_padded
    for i, (name, lay) in enumerate(cl_sequence_padded().items()):
        _rollout_compare(lay, random_reset=bool(i % 2), max_steps=29, N=515, T=64, seed=40 + i, part=part, every=4,
                         log=True)


@pytest.mark.parametrize("N", [1, 2, 15, 16, 17, 31, 33, 64, 1000])
def test_ragged_batch_sizes(N):
    from jaxovercooked_b200.layouts import overcooked_layouts
    _rollout_compare(overcooked_layouts["cramped_room"], random_reset=True, max_steps=11, N=N, T=25, seed=N)
    _rollout_compare(overcooked_layouts["coord_ring"], random_reset=False, max_steps=11, N=N, T=14, seed=N + 1)


def test_functional_step_leaves_input_untouched():
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["asymm_advantages"]
    env = ocb.make("overcooked", layout=lay, random_reset=True, max_steps=9)
    keys = ocr.split(ocr.PRNGKey(3), 100)
    obs, st = env.reset(keys)
    before = gh.state_to_numpy(st)
    a = torch.randint(0, 6, (100,), dtype=torch.int32, device="cuda")
    obs2, st2, *_ = env.step(ocr.split(ocr.PRNGKey(4), 100), st, {"agent_0": a, "agent_1": a})
    after = gh.state_to_numpy(st)
    for f in before:
        assert_same(after[f], before[f], f"input State.{f} after functional step")
    assert st2.maze_map.data_ptr() != st.maze_map.data_ptr()
    # donated step from the same input gives the same result as the functional one
    obs3, st3, *_ = env.step(ocr.split(ocr.PRNGKey(4), 100), st.clone(), {"agent_0": a, "agent_1": a}, donate=True)
    gh.compare_state(st3, gh.state_to_numpy(st2), "donated vs functional")
    assert torch.equal(obs3["agent_0"], obs2["agent_0"]) and torch.equal(obs3["agent_1"], obs2["agent_1"])
    # get_obs reproduces the step's observation
    o = env.get_obs(st2)
    assert torch.equal(o["agent_0"], obs2["agent_0"]) and torch.equal(o["agent_1"], obs2["agent_1"])


def test_unbatched_call():
    """cl_methods/EWC.py:133,144 and IPPO_CL.py:988 call env.step without vmap: batch shape ()."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["cramped_room"]
    env = ocb.make("overcooked", layout=lay)
    ref = orc.Oracle(lay)
    key = ocr.PRNGKey(5)
    obs, st = env.reset(key)
    assert obs["agent_0"].shape == (4, 5, 26) and st.maze_map.shape == (12, 13, 3) and st.time.shape == ()
    (r0, r1), rst = ref.reset(key.cpu().numpy()[None])
    assert_same(obs["agent_0"].cpu().numpy(), r0[0], "obs0")
    for t in range(5):
        k = ocr.split(ocr.PRNGKey(t), 2)[1]
        act = {"agent_0": torch.tensor(5, dtype=torch.int32, device="cuda"), "agent_1": torch.tensor(3, dtype=torch.int32, device="cuda")}
        obs, st, rew, done, info = env.step(k, st, act)
        (r0, r1), rr, s0, s1, dd = ref.step(k.cpu().numpy()[None], rst, np.array([5], np.int32), np.array([3], np.int32))
        assert_same(obs["agent_1"].cpu().numpy(), r1[0], f"t={t} obs1")
        assert rew["agent_0"].shape == () and done["__all__"].shape == ()


def test_reset_state_argument():
    """multi_agent_env.py:55-59: a caller-supplied reset_state replaces reset(key_reset) when done."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["coord_ring"]
    env = ocb.make("overcooked", layout=lay, max_steps=3, random_reset=True)
    N = 50
    _, rs = env.reset(ocr.split(ocr.PRNGKey(11), N))
    obs, st = env.reset(ocr.split(ocr.PRNGKey(12), N))
    a = {"agent_0": torch.full((N,), 4, dtype=torch.int32, device="cuda"), "agent_1": torch.full((N,), 4, dtype=torch.int32, device="cuda")}
    for t in range(3):
        obs, st, rew, done, info = env.step(ocr.split(ocr.PRNGKey(20 + t), N), st, a, reset_state=rs)
    assert bool(done["__all__"].all())
    gh.compare_state(st, gh.state_to_numpy(rs), "state after done == reset_state", fields=DYN_FIELDS)
    o = env.get_obs(rs)
    assert torch.equal(obs["agent_0"], o["agent_0"]) and torch.equal(obs["agent_1"], o["agent_1"])


def test_prng_split_matches_oracle():
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    from oracle import prng
    for impl, part in (("threefry_legacy", False), ("threefry_partitionable", True)):
        for num in (1, 2, 3, 16, 1001, 65536):
            key = ocr.PRNGKey(num)
            got = ocr.split(key, num, impl=impl).cpu().numpy()
            assert_same(got, prng.split(prng.PRNGKey(num), num, part), f"split {impl} {num}")
        # shard view: rows [first, first+count) of a global split
        full = prng.split(prng.PRNGKey(9), 4096, part)
        got = ocr.split(ocr.PRNGKey(9), 4096, first=1024, count=512, impl=impl).cpu().numpy()
        assert_same(got, full[1024:1536], f"sharded split {impl}")


def test_lazy_split_equals_materialised_keys():
    """keys=NULL + split_key: the kernel derives split(key, num)[first+i] itself, only for envs that reset."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["forced_coord"]
    for impl in ("threefry_legacy", "threefry_partitionable"):
        env = ocb.make("overcooked", layout=lay, random_reset=True, max_steps=5, prng_impl=impl)
        num, first, N = 5000, 1234, 999
        _, st_a = env.reset(ocr.split(ocr.PRNGKey(1), num, first=first, count=N, impl=impl))
        st_b = st_a.clone()
        rng = ocr.PRNGKey(2)
        subs = ocr.split_chain(rng, 12, impl=impl)
        ref_rng = ocr.PRNGKey(2)
        for t in range(12):
            ks = ocr.split(ref_rng, 2, impl=impl); ref_rng = ks[0]
            assert torch.equal(ks[1].view(torch.int32), subs[t].view(torch.int32))
            a = {"agent_0": torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda"),
                 "agent_1": torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda")}
            oa, st_a, *_ = env.step(ocr.split(subs[t], num, first=first, count=N, impl=impl), st_a, a, donate=True)
            ob, st_b, *_ = env.step(ocr.LazySplit(subs[t], num, first, N), st_b, a, donate=True)
            assert torch.equal(oa["agent_0"], ob["agent_0"]) and torch.equal(oa["agent_1"], ob["agent_1"])
            gh.compare_state(st_b, gh.state_to_numpy(st_a), f"{impl} t={t}")
        assert torch.equal(rng.view(torch.int32), ref_rng.view(torch.int32))


def test_stats_accumulators_exact():
    ocb, orc, gh = _mods()
    lay = ocb.overcooked_layouts["cramped_room"]
    env = ocb.LogWrapper(ocb.make("overcooked", layout=lay, random_reset=True, max_steps=13))
    ref = orc.Oracle(lay, random_reset=True, max_steps=13)
    N, T = 777, 40
    rng = np.random.default_rng(5)
    rk = rng.integers(0, 2**32, (N, 2), dtype=np.uint32)
    obs, st = env.reset(gh.keys_to_dev(rk))
    _, rst = ref.reset(rk)
    rlog = ref.new_log(N)
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    exp = np.zeros(8, np.int64)
    for t in range(T):
        sk = rng.integers(0, 2**32, (N, 2), dtype=np.uint32)
        a = rng.choice(6, size=(N, 2), p=ACTION_P).astype(np.int32)
        obs, st, rew, done, info = env.step(gh.keys_to_dev(sk), st, {"agent_0": gh.to_dev(a[:, 0]), "agent_1": gh.to_dev(a[:, 1])},
                                            donate=True, stats=stats)
        _, rr, s0, s1, dd = ref.step(sk, rst, a[:, 0], a[:, 1], log=rlog)
        exp[0] += dd.sum(); exp[1] += int(rr.sum()); exp[2] += int(s0.sum()); exp[3] += int(s1.sum())
        exp[4] += int(rlog["returned_episode_returns"][dd, 0].sum()); exp[5] += int(rlog["returned_episode_lengths"][dd, 0].sum())
        exp[6] += N
    assert stats.cpu().numpy().tolist() == exp.tolist()


def test_full_size_cramped_room_65536():
    """BASELINE.json config 2: cramped_room, 65,536 envs, crossing the auto-reset at step 400.
    All envs are compared at t in {1, 399, 400, 401}; a 1,024-env prefix every 8th step; invariants at the end."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    from oracle import prng
    lay = ocb.overcooked_layouts["cramped_room"]
    N, T = 65536, 404
    env = ocb.make("overcooked", layout=lay)
    ref = orc.Oracle(lay)
    rng = ocr.PRNGKey(0)
    rng_np = prng.PRNGKey(0)
    arng = np.random.default_rng(0)
    ks = ocr.split(rng, 2); rng, sub = ks[0], ks[1]
    obs, st = env.reset(ocr.split(sub, N))
    kn = prng.split(rng_np); rng_np, subn = kn[0], kn[1]
    (r0, r1), rst = ref.reset(prng.split(subn, N))
    assert_same(obs["agent_0"].cpu().numpy(), r0, "reset obs0")
    out = ref.alloc_out(N)
    for t in range(1, T + 1):
        ks = ocr.split(rng, 2); rng, sub = ks[0], ks[1]
        keys = ocr.split(sub, N)
        a = arng.integers(0, 6, (N, 2)).astype(np.int32)
        obs, st, rew, done, info = env.step(keys, st, {"agent_0": gh.to_dev(a[:, 0]), "agent_1": gh.to_dev(a[:, 1])}, donate=True)
        kn = prng.split(rng_np); rng_np, subn = kn[0], kn[1]
        (r0, r1), rr, s0, s1, dd = ref.step(prng.split(subn, N), rst, a[:, 0], a[:, 1], out=out)
        if t in (1, 399, 400, 401):
            assert_same(obs["agent_0"].cpu().numpy(), r0, f"t={t} obs0 (all envs)")
            assert_same(obs["agent_1"].cpu().numpy(), r1, f"t={t} obs1 (all envs)")
            gh.compare_state(st, rst, f"t={t} (all envs)")
            assert_same(done["__all__"].cpu().numpy(), dd, f"t={t} done")
        elif t % 8 == 0:
            assert_same(obs["agent_0"][:1024].cpu().numpy(), r0[:1024], f"t={t} obs0 (prefix)")
            assert_same(obs["agent_1"][:1024].cpu().numpy(), r1[:1024], f"t={t} obs1 (prefix)")
            assert_same(st.maze_map[:1024].cpu().numpy(), rst["maze_map"][:1024], f"t={t} maze_map (prefix)")
        assert_same(rew["agent_0"].cpu().numpy(), rr, f"t={t} reward")
    # SURVEY App. A.8 invariants on the final state
    mm = st.maze_map.cpu().numpy()
    ring = np.ones(mm.shape[1:3], bool); ring[4:-4, 4:-4] = False
    assert (mm[:, ring] == np.array([2, 5, 0], np.uint8)).all()
    assert ((mm[..., 0] == 10).sum(axis=(1, 2)) == 2).all()
    assert np.isin(st.agent_inv.cpu().numpy(), [1, 3, 5, 9]).all()
    o0 = obs["agent_0"].cpu().numpy()
    assert (o0[..., 0].sum(axis=(1, 2)) == 1).all() and (o0[..., 1].sum(axis=(1, 2)) == 1).all()
    assert (o0[..., [13, 17, 19, 24]] == 0).all() and o0.max() <= 19


def _big_run(lay, *, N, T, random_reset, seed, full_at, prefix=256, prefix_every=16, max_steps=400):
    """Caller recipe (IPPO_CL.py:477,533-538) at benchmark sizes: all envs compared at `full_at`, a prefix often."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    from oracle import prng
    env = ocb.make("overcooked", layout=lay, random_reset=random_reset, max_steps=max_steps)
    ref = orc.Oracle(lay, random_reset=random_reset, max_steps=max_steps)
    rng, rng_np = ocr.PRNGKey(seed), prng.PRNGKey(seed)
    arng = np.random.default_rng(seed)
    ks = ocr.split(rng, 2); rng, sub = ks[0].clone(), ks[1]
    obs, st = env.reset(ocr.split(sub, N))
    kn = prng.split(rng_np); rng_np, subn = kn[0], kn[1]
    (r0, r1), rst = ref.reset(prng.split(subn, N))
    assert_same(obs["agent_0"].cpu().numpy(), r0, "reset obs0")
    gh.compare_state(st, rst, "reset")
    out = ref.alloc_out(N)
    subs = ocr.split_chain(rng, T)                       # the whole `rng, _rng = split(rng)` chain in one kernel
    for t in range(1, T + 1):
        a = arng.integers(0, 6, (N, 2)).astype(np.int32)
        obs, st, rew, done, info = env.step(ocr.LazySplit(subs[t - 1], N), st,
                                            {"agent_0": gh.to_dev(a[:, 0]), "agent_1": gh.to_dev(a[:, 1])}, donate=True)
        kn = prng.split(rng_np); rng_np, subn = kn[0], kn[1]
        (r0, r1), rr, s0, s1, dd = ref.step(prng.split(subn, N), rst, a[:, 0], a[:, 1], out=out)
        if t in full_at:
            assert_same(obs["agent_0"].cpu().numpy(), r0, f"t={t} obs0 (all envs)")
            assert_same(obs["agent_1"].cpu().numpy(), r1, f"t={t} obs1 (all envs)")
            gh.compare_state(st, rst, f"t={t} (all envs)")
            assert_same(done["__all__"].cpu().numpy(), dd, f"t={t} done")
            assert_same(info["shaped_reward"]["agent_0"].cpu().numpy(), s0, f"t={t} shaped0")
        elif t % prefix_every == 0:
            assert_same(obs["agent_0"][:prefix].cpu().numpy(), r0[:prefix], f"t={t} obs0 (prefix)")
            assert_same(obs["agent_1"][:prefix].cpu().numpy(), r1[:prefix], f"t={t} obs1 (prefix)")
            assert_same(st.maze_map[:prefix].cpu().numpy(), rst["maze_map"][:prefix], f"t={t} maze_map (prefix)")
        assert_same(rew["agent_0"].cpu().numpy(), rr, f"t={t} reward")


def test_config1_16_envs_400_steps_logwrapper():
    """BASELINE.json configs[0]: cramped_room, 16 envs, LogWrapper, 400-step episodes, seed 30."""
    from jaxovercooked_b200.layouts import overcooked_layouts
    _rollout_compare(overcooked_layouts["cramped_room"], random_reset=False, max_steps=400, N=16, T=402, seed=30,
                     every=1, log=True)


@pytest.mark.parametrize("idx", range(5))
def test_config3_cl_sequence_32768_envs(idx):
    """BASELINE.json configs[2], one GPU's share: each padded 5x9 CL layout, 32,768 envs, across the auto-reset."""
    from jaxovercooked_b200.layouts import cl_sequence_padded
    name, lay = list(cl_sequence_padded().items())[idx]
    _big_run(lay, N=32768, T=404, random_reset=False, seed=idx, full_at=(1, 399, 400, 401, 404))


def test_config5_stress_10x10_random_reset_65536():
    """BASELINE.json configs[4]: padded 10x10 layout, random_reset=True, 65,536 envs, across the auto-reset."""
    from jaxovercooked_b200.layouts import overcooked_layouts
    _big_run(overcooked_layouts["easy_layout_padded"], N=65536, T=70, random_reset=True, seed=9,
             full_at=(1, 59, 60, 61, 70), max_steps=60)


def test_config5_stress_9x11_random_reset():
    from jaxovercooked_b200.layouts import overcooked_layouts
    _big_run(overcooked_layouts["bottleneck_large"], N=16384 + 37, T=45, random_reset=True, seed=10,
             full_at=(1, 29, 30, 31, 45), max_steps=30)


def test_step_host_matches_device_step():
    """oc_step_host (host action / result buffers) gives exactly what oc_step gives."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["counter_circuit"]
    env = ocb.make("overcooked", layout=lay, random_reset=True, max_steps=7)
    N = 1000
    _, st_a = env.reset(ocr.split(ocr.PRNGKey(1), N))
    st_b = st_a.clone()
    hs = env.host_stepper(st_b)
    subs = ocr.split_chain(ocr.PRNGKey(2), 20)
    torch.cuda.synchronize()
    for t in range(20):
        a = torch.randint(0, 6, (2, N), dtype=torch.int32)
        hs.actions.copy_(a)
        hs.step(subs[t])
        d = a.cuda()
        obs, st_a, rew, done, info = env.step(ocr.LazySplit(subs[t], N), st_a, {"agent_0": d[0], "agent_1": d[1]}, donate=True)
        hs.wait()
        torch.cuda.synchronize()
        assert torch.equal(hs.obs["agent_0"], obs["agent_0"]) and torch.equal(hs.obs["agent_1"], obs["agent_1"])
        assert torch.equal(hs.reward, rew["agent_0"].cpu()) and torch.equal(hs.done, done["__all__"].cpu())
        assert torch.equal(hs.shaped0, info["shaped_reward"]["agent_0"].cpu())
        assert torch.equal(hs.shaped1, info["shaped_reward"]["agent_1"].cpu())
        gh.compare_state(st_b, gh.state_to_numpy(st_a), f"t={t}")


def test_rollout_buffer_is_batchify_layout():
    """the step writes observations straight into the (T, 2N, h*w*26) agent-major trajectory tensor
    that baselines/utils.py:48-61 `batchify` + `Transition.obs` build with an extra copy."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["cramped_room"]
    env = ocb.make("overcooked", layout=lay, max_steps=9)
    N, T = 512, 12
    buf = ocb.RolloutBuffer(env, T, N)
    _, st = env.reset(ocr.split(ocr.PRNGKey(0), N))
    st_ref = st.clone()
    subs = ocr.split_chain(ocr.PRNGKey(1), T)
    for t in range(T):
        a = {"agent_0": torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda"),
             "agent_1": torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda")}
        _, st, _, _, _ = env.step(ocr.LazySplit(subs[t], N), st, a, donate=True, out=buf.slot(t))
        obs, st_ref, rew, done, info = env.step(ocr.LazySplit(subs[t], N), st_ref, a, donate=True)
        batchified = torch.stack([obs["agent_0"], obs["agent_1"]]).reshape(2 * N, -1)      # utils.py:48-61
        assert torch.equal(buf.batchified_obs(t), batchified)
        assert torch.equal(buf.reward[t], rew["agent_0"]) and torch.equal(buf.done[t], done["__all__"])
        assert torch.equal(buf.shaped[t, 1], info["shaped_reward"]["agent_1"])
    assert buf.batchified_obs().shape == (T, 2 * N, 4 * 5 * 26)


def test_ct_rollout_manager_vdn_surface():
    """CTRolloutManager(LogWrapper(env), batch_size, preprocess_obs=False) as vdn_mlp.py:339-346 uses it."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    from oracle import prng
    lay = ocb.overcooked_layouts["cramped_room"]
    B = 300
    mgr = ocb.CTRolloutManager(ocb.LogWrapper(ocb.make("overcooked", layout=lay, max_steps=6)), batch_size=B)
    ref = orc.Oracle(lay, max_steps=6)
    k0 = ocr.PRNGKey(3)
    obs, st = mgr.batch_reset(k0)
    (r0, r1), rst = ref.reset(prng.split(prng.PRNGKey(3), B))
    assert obs["__all__"].shape == (B, 2 * 4 * 5 * 26)
    assert_same(obs["__all__"].cpu().numpy(), np.concatenate([r0.reshape(B, -1), r1.reshape(B, -1)], 1), "reset __all__")
    rlog = ref.new_log(B)
    for t in range(9):
        a = np.random.default_rng(t).integers(0, 6, (B, 2)).astype(np.int32)
        kt = ocr.PRNGKey(100 + t)
        obs, st, rew, done, info = mgr.batch_step(kt, st, {"agent_0": gh.to_dev(a[:, 0]), "agent_1": gh.to_dev(a[:, 1])})
        (r0, r1), rr, s0, s1, dd = ref.step(prng.split(prng.PRNGKey(100 + t), B), rst, a[:, 0], a[:, 1], log=rlog)
        assert_same(obs["__all__"].cpu().numpy(), np.concatenate([r0.reshape(B, -1), r1.reshape(B, -1)], 1), f"t={t} __all__")
        assert_same(rew["__all__"].cpu().numpy(), rr, f"t={t} reward __all__")
        assert_same(info["returned_episode_returns"].cpu().numpy(), rlog["returned_episode_returns"], f"t={t} returns")


def _big_layout(h, w):
    """Enclosed h x w kitchen with everything on the walls and a few interior counters (max-size stress)."""
    rows = [["W"] * w for _ in range(h)]
    for y in range(1, h - 1):
        for x in range(1, w - 1):
            rows[y][x] = " " if (x % 4 or y % 3) else "W"
    rows[0][2], rows[0][w - 3] = "P", "P"
    rows[h - 1][2], rows[h - 1][w - 3] = "O", "X"
    rows[h // 2][0], rows[h // 2][w - 1] = "B", "O"
    rows[1][1], rows[h - 2][w - 2] = "A", "A"
    from jaxovercooked_b200.layouts import layout_grid_to_dict
    return layout_grid_to_dict("\n".join("".join(r) for r in rows))


@pytest.mark.parametrize("h,w", [(15, 15), (16, 16), (3, 16), (16, 3)])
def test_maximum_grid_sizes(h, w):
    """env_generator.py:220-223 allows up to 15x15; the ABI accepts up to 16x16."""
    _rollout_compare(_big_layout(h, w), random_reset=True, max_steps=13, N=77, T=30, seed=h * w, every=2)


def test_unaligned_observation_buffers_take_the_plain_store_path():
    """Observation pointers that are not 16-byte aligned cannot use the bulk-copy engine; results must not change."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    lay = ocb.overcooked_layouts["cramped_room"]
    env = ocb.make("overcooked", layout=lay, random_reset=True, max_steps=6)
    N = 100
    _, st_a = env.reset(ocr.split(ocr.PRNGKey(0), N))
    st_b = st_a.clone()
    nbytes = N * 4 * 5 * 26
    raw = torch.zeros(2 * nbytes + 64, dtype=torch.uint8, device="cuda")
    o0 = raw[2:2 + nbytes].view(N, 4, 5, 26)                       # 2-byte aligned only
    o1 = raw[nbytes + 22:2 * nbytes + 22].view(N, 4, 5, 26)
    subs = ocr.split_chain(ocr.PRNGKey(1), 10)
    for t in range(10):
        a = {"agent_0": torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda"),
             "agent_1": torch.randint(0, 6, (N,), dtype=torch.int32, device="cuda")}
        out = {"obs0": o0, "obs1": o1, "reward": torch.empty(N, device="cuda"), "shaped0": torch.empty(N, device="cuda"),
               "shaped1": torch.empty(N, device="cuda"), "done": torch.empty(N, dtype=torch.bool, device="cuda")}
        _, st_a, _, _, _ = env.step(ocr.LazySplit(subs[t], N), st_a, a, donate=True, out=out)
        obs, st_b, *_ = env.step(ocr.LazySplit(subs[t], N), st_b, a, donate=True)
        assert torch.equal(o0, obs["agent_0"]) and torch.equal(o1, obs["agent_1"])


def test_caller_buffers_are_validated_before_they_cross_the_abi():
    """Raw pointers go to the kernels: wrong dtype / shape / device / non-contiguous buffers must raise, not corrupt memory."""
    ocb, orc, gh = _mods()
    from jaxovercooked_b200 import random as ocr
    env = ocb.make("overcooked", layout=ocb.overcooked_layouts["cramped_room"])
    N = 12
    keys = ocr.split(ocr.PRNGKey(0), N)
    _, st = env.reset(keys)
    act = {"agent_0": torch.zeros(N, dtype=torch.int32, device="cuda"), "agent_1": torch.zeros(N, dtype=torch.int32, device="cuda")}

    def good():
        return {"obs0": torch.empty((N, 4, 5, 26), dtype=torch.uint8, device="cuda"),
                "obs1": torch.empty((N, 4, 5, 26), dtype=torch.uint8, device="cuda"),
                "reward": torch.empty(N, dtype=torch.float32, device="cuda"), "shaped0": torch.empty(N, dtype=torch.float32, device="cuda"),
                "shaped1": torch.empty(N, dtype=torch.float32, device="cuda"), "done": torch.empty(N, dtype=torch.bool, device="cuda")}
    env.step(keys, st, act, out=good())
    bad = [("reward", torch.empty(N, dtype=torch.float64, device="cuda")),
           ("obs0", torch.empty((N, 4, 5, 26), dtype=torch.bool, device="cuda")),
           ("obs1", torch.empty((N, 5, 4, 26), dtype=torch.uint8, device="cuda")),
           ("obs0", torch.empty((2 * N, 4, 5, 26), dtype=torch.uint8, device="cuda")[::2]),
           ("done", torch.empty(N, dtype=torch.int32, device="cuda")),
           ("shaped1", torch.empty(N, dtype=torch.float32)),
           ("shaped0", torch.empty(N + 1, dtype=torch.float32, device="cuda"))]
    for k, t in bad:
        o = good()
        o[k] = t
        with pytest.raises(ValueError):
            env.step(keys, st, act, out=o)
    z = lambda: torch.zeros((N, 2), dtype=torch.float32, device="cuda")
    log = {"episode_returns": z(), "episode_lengths": z(), "returned_episode_returns": z(), "returned_episode_lengths": z()}
    env.step(keys, st, act, log=dict(log), stats=torch.zeros(8, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        env.step(keys, st, act, log=dict(log, episode_lengths=torch.zeros(N, dtype=torch.float32, device="cuda")))
    with pytest.raises(ValueError):
        env.step(keys, st, act, log={k: v for k, v in log.items() if k != "episode_returns"})
    with pytest.raises(ValueError):
        env.step(keys, st, act, stats=torch.zeros(8, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        env.step(keys, st, act, stats=torch.zeros(4, dtype=torch.int64, device="cuda"))
    hs = env.host_stepper(st)
    with pytest.raises(ValueError):
        hs.step(ocr.PRNGKey(1), torch.zeros((2, N), dtype=torch.int32))           # not pinned
    with pytest.raises(ValueError):
        hs.step(ocr.PRNGKey(1), torch.zeros((N, 2), dtype=torch.int32).pin_memory())
    hs.step(ocr.PRNGKey(1), torch.zeros((2, N), dtype=torch.int32).pin_memory())
    hs.wait()
