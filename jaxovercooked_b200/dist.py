"""Multi-GPU layout of the path: environments shard trivially, one process per GPU, no per-step communication.

GPU `rank` of `world` owns the contiguous slice [first, first+count) of the global env axis.  Per-env keys are
derived from the GLOBAL env index (`random.split(key, n_global, first=..., count=...)` or `random.LazySplit`),
so a trajectory is bit-identical for any number of GPUs.  The only collective is one all-reduce (sum) of the
int64[8] statistics vector that `oc_step` accumulates (include/oc_b200.h, oc_step_extras.stats) -- all addends are
integers, so the result is order-independent.  It stands in for the `tree_map(mean)` over rollout `info` that the
reference's trainers do on one device (baselines/IPPO_CL.py:771-812).
"""
import os

STAT_NAMES = ("finished_episodes", "sum_reward", "sum_shaped_agent_0", "sum_shaped_agent_1",
              "sum_returned_episode_returns", "sum_returned_episode_lengths", "env_steps", "reserved")


def world_info():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_global, rank, world):
    """Contiguous, balanced partition of range(n_global): returns (first, count) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    base, rem = divmod(int(n_global), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def reduce_stats(stats, group=None):
    """In-place sum of the int64[8] statistics vector over all ranks (NCCL for CUDA tensors, gloo for CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def stats_to_metrics(stats):
    """Means the trainers log, formed AFTER the reduction from exact integer sums."""
    v = [int(x) for x in stats.tolist()]
    d = dict(zip(STAT_NAMES, v))
    steps = max(d["env_steps"], 1)
    eps = max(d["finished_episodes"], 1)
    return {"env_steps": d["env_steps"], "finished_episodes": d["finished_episodes"],
            "reward_mean": d["sum_reward"] / steps,
            "shaped_reward_mean/agent_0": d["sum_shaped_agent_0"] / steps,
            "shaped_reward_mean/agent_1": d["sum_shaped_agent_1"] / steps,
            "returned_episode_returns_mean": d["sum_returned_episode_returns"] / eps,
            "returned_episode_lengths_mean": d["sum_returned_episode_lengths"] / eps}
