"""The N > 1 host path on CPU: two gloo ranks each own a shard of the env axis, derive their keys from GLOBAL env
indices, accumulate the integer statistics vector (here with the CPU oracle standing in for the kernel) and
all-reduce it.  The result must equal the single-process run over the whole batch, and trajectories must not
depend on the world size."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxovercooked_b200 import dist as ocd
from jaxovercooked_b200.layouts import overcooked_layouts


def test_shard_range_partitions_exactly():
    for n in (1, 2, 7, 16, 65536, 262144, 1000003):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                first, count = ocd.shard_range(n, r, world)
                assert first == seen and count >= 0
                seen += count
            assert seen == n
    with pytest.raises(ValueError):
        ocd.shard_range(10, 2, 2)


def _rollout_stats(first, count, n_global, T, seed):
    """Statistics of envs [first, first+count) of a global batch, keys by global index (oracle as the stepper)."""
    from oracle import oracle as orc
    from oracle import prng
    lay = overcooked_layouts["cramped_room"]
    o = orc.Oracle(lay, random_reset=True, max_steps=11)
    rng = prng.PRNGKey(seed)
    rng, sub = prng.split(rng)
    _, st = o.reset(prng.split(sub, n_global)[first:first + count])
    log = o.new_log(count)
    acts = np.random.default_rng(seed).integers(0, 6, (T, n_global, 2)).astype(np.int32)
    stats = np.zeros(8, np.int64)
    sig = 0
    for t in range(T):
        rng, sub = prng.split(rng)
        keys = prng.split(sub, n_global)[first:first + count]
        a = acts[t, first:first + count]
        (o0, o1), rew, s0, s1, done = o.step(keys, st, a[:, 0].copy(), a[:, 1].copy(), log=log)
        stats[0] += done.sum(); stats[1] += int(rew.sum()); stats[2] += int(s0.sum()); stats[3] += int(s1.sum())
        stats[4] += int(log["returned_episode_returns"][done, 0].sum())
        stats[5] += int(log["returned_episode_lengths"][done, 0].sum())
        stats[6] += count
        sig += int(o0.astype(np.int64).sum()) * (t + 1) + int(st["agent_pos"].astype(np.int64).sum())
    return stats, sig


def _worker(rank, world, port, n_global, T, seed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = ocd.shard_range(n_global, rank, world)
        stats, sig = _rollout_stats(first, count, n_global, T, seed)
        t = torch.from_numpy(stats.copy())
        sg = torch.tensor([sig], dtype=torch.int64)
        ocd.reduce_stats(t)
        dist.all_reduce(sg)
        if rank == 0:
            out.put((t.tolist(), int(sg.item())))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_two_rank_gloo_equals_single_process():
    n_global, T, seed = 203, 30, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, T, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_stats, got_sig = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_stats, ref_sig = _rollout_stats(0, n_global, n_global, T, seed)
    assert got_stats == ref_stats.tolist()
    assert got_sig == ref_sig          # same trajectories regardless of the number of ranks
    m = ocd.stats_to_metrics(torch.tensor(got_stats))
    assert m["env_steps"] == n_global * T and m["finished_episodes"] == ref_stats[0]
