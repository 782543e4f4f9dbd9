#!/usr/bin/env python
"""Policy forward against the float32 oracle on rows from a real rollout: max / mean abs error of logits and value, and how
often the sampled action differs (same key).  Run once per build (OC_B200_LIB selects a variant library)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import jaxovercooked_b200 as ocb
from jaxovercooked_b200 import random as ocr
from jaxovercooked_b200.policy import ActorCritic
from oracle import policy_oracle as po
N, T = 4096, 60
env = ocb.make("overcooked", layout=ocb.overcooked_layouts["cramped_room"], random_reset=True)
_, st = env.reset(ocr.split(ocr.PRNGKey(0), N))
g = torch.Generator(device="cuda").manual_seed(1)
acts = {a: torch.randint(0, 6, (T, N), dtype=torch.int32, device="cuda", generator=g) for a in ("agent_0", "agent_1")}
env.rollout(ocr.split_chain(ocr.PRNGKey(1), T), st, acts)
for scale, name in ((1.0, "orthogonal init"), (8.0, "hidden kernels x8 (saturating tanh)")):
    params = po.make_params(3, env.height * env.width * 26)
    if scale != 1.0:
        for k in ("Dense_0", "Dense_1", "Dense_3", "Dense_4"):
            params["params"][k]["kernel"] = params["params"][k]["kernel"] * scale
    net = ActorCritic(6)
    P = net.params_from_numpy(params)
    net.bind(P, env=env)
    cells = env.empty_cells(N)
    obs = env.get_obs(st, cells=cells)
    action, log_prob, value, logits = net.act_cells(P, cells, ocr.PRNGKey(7), want_logits=True)
    rows = np.concatenate([obs["agent_0"].cpu().numpy().reshape(N, -1), obs["agent_1"].cpu().numpy().reshape(N, -1)])
    lo, vo = po.forward(params, rows)
    dl, dv = np.abs(logits.cpu().numpy() - lo), np.abs(value.cpu().numpy() - vo)
    print(json.dumps({"lib": os.path.basename(os.environ.get("OC_B200_LIB", "in-tree")), "params": name, "rows": 2 * N,
                      "logits_max_abs_err": float(dl.max()), "logits_mean_abs_err": float(dl.mean()), "logits_max_abs": float(np.abs(lo).max()),
                      "value_max_abs_err": float(dv.max()), "value_mean_abs_err": float(dv.mean()), "value_max_abs": float(np.abs(vo).max())}), flush=True)
