"""BASELINE.json configs[3] shape: the rollout phase of an IPPO update (4,096 envs x 128 steps) with the actor-critic policy
in the loop, as one CUDA graph (jaxovercooked_b200.PolicyRollout).  usage: rollout_config4.py [--envs 4096] [--steps 128]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import jaxovercooked_b200 as ocb                                     # noqa: E402
from jaxovercooked_b200 import random as jr                          # noqa: E402
from jaxovercooked_b200.policy import ActorCritic                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--layout", default="cramped_room")
    ap.add_argument("--rollouts", type=int, default=20)
    a = ap.parse_args()
    env = ocb.make("overcooked", layout=ocb.overcooked_layouts[a.layout])
    net = ActorCritic(env.action_space().n)
    params = net.init(0, env.height * env.width * 26)
    with torch.no_grad():
        params["params"]["Dense_2"]["kernel"].mul_(100.0)
    rng = jr.PRNGKey(30)
    rng, sub = jr.split(rng)
    _, state = env.reset(jr.split(sub, a.envs))
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    ro = ocb.PolicyRollout(env, net, params, state, rng, a.steps, stats=stats)
    for _ in range(3):
        ro.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.rollouts):
        ro.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.rollouts
    print(json.dumps({"envs": a.envs, "steps": a.steps, "layout": a.layout, "ms_per_rollout": round(ms, 3),
                      "us_per_step": round(ms * 1e3 / a.steps, 2), "env_steps_per_s_M": round(a.envs * a.steps / ms / 1e3, 1),
                      "kernels_per_rollout": 2 * a.steps + 1, "stats": stats.tolist()}))


if __name__ == "__main__":
    main()
