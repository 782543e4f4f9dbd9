"""Timing probe for oc_policy_forward: rows from a real rollout (both agents of N envs), CUDA events, rotating obs
slots so that the rows come from HBM.  Usage: python profiles/probes/policy_bench.py [--envs 65536] [--layout cramped_room]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import jaxovercooked_b200 as ocb                                     # noqa: E402
from jaxovercooked_b200 import random as jr                          # noqa: E402
from jaxovercooked_b200.policy import ActorCritic                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--layout", default="cramped_room")
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--warm-steps", type=int, default=30)
    a = ap.parse_args()
    env = ocb.make("overcooked", layout=ocb.overcooked_layouts[a.layout])
    n, K = a.envs, env.height * env.width * 26
    net = ActorCritic(6)
    params = net.init(0, K)
    with torch.no_grad():
        params["params"]["Dense_2"]["kernel"].mul_(100.0)
    rng = jr.PRNGKey(0)
    rng, sub = jr.split(rng)
    obs, state = env.reset(jr.split(sub, n))
    # walk the envs away from the reset state with the policy itself, then record `slots` consecutive obs batches
    slots = torch.empty((a.slots, 2 * n, K), dtype=torch.uint8, device="cuda")
    cslots = torch.empty((a.slots, n, env.info.cell_stride), dtype=torch.uint8, device="cuda")
    cells = env.empty_cells(n)
    env.get_obs(state, cells=cells)
    outb = {"obs0": torch.empty_like(obs["agent_0"]), "obs1": torch.empty_like(obs["agent_1"]),
            "reward": torch.empty(n, device="cuda"), "shaped0": torch.empty(n, device="cuda"),
            "shaped1": torch.empty(n, device="cuda"), "done": torch.empty(n, dtype=torch.bool, device="cuda"), "cells": cells}
    for t in range(a.warm_steps + a.slots):
        rng, k_act, k_step = jr.split(rng, 3)
        batch = torch.cat([obs["agent_0"].reshape(n, -1), obs["agent_1"].reshape(n, -1)], dim=0)
        if t >= a.warm_steps:
            slots[t - a.warm_steps].copy_(batch)
            cslots[t - a.warm_steps].copy_(cells)
        action, _, _ = net.act(params, batch, k_act)
        obs, state, *_ = env.step(jr.split(k_step, n), state, {"agent_0": action[:n], "agent_1": action[n:]}, donate=True, out=outb)
    tm = env.obs_template().reshape(-1)
    dev = (slots[0] != torch.from_numpy(tm).cuda()).sum(1).float()
    res = {"envs": n, "rows": 2 * n, "K": K, "deviations_per_row": round(float(dev.mean()), 2), "max_dev": int(dev.max())}
    out = {k: torch.empty((2 * n,) + s, dtype=d, device="cuda") for k, s, d in
           (("logits", (6,), torch.float32), ("value", (), torch.float32), ("action", (), torch.int32), ("log_prob", (), torch.float32))}
    for name, tmpl, key in (("cells+sample", "cells", rng), ("cells", "cells", None), ("template+sample", tm, rng), ("template", tm, None),
                            ("no_template+sample", None, rng)):
        if isinstance(tmpl, str):
            net.bind(params, env=env)
        else:
            net.bind(params, obs_template=tmpl)
        def run(i):
            if isinstance(tmpl, str):
                net._forward_cells(params, cslots[i % a.slots], key, out=out)
            elif key is None:
                net._forward(params, slots[i % a.slots], None, out=out)
            else:
                net._forward(params, slots[i % a.slots], key, out=out)
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        res[name] = {"us": round(us, 2), "rows_per_s_G": round(2 * n / us / 1e3, 3), "obs_GBps": round(2 * n * K / us / 1e3, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
