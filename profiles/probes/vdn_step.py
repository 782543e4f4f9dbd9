#!/usr/bin/env python
"""CTRolloutManager.batch_step (the VDN baselines' env step): obs["__all__"] written by the kernel (obs_interleaved) against the
two-array step followed by torch.cat (round 1's path).  us per step in a replayed CUDA graph of 16 steps, rotating buffers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import jaxovercooked_b200 as ocb
from jaxovercooked_b200 import random as ocr

dev = torch.device("cuda")
for N in (4096, 65536):
    lay = ocb.overcooked_layouts["cramped_room"]
    env = ocb.make("overcooked", layout=lay)
    mgr = ocb.CTRolloutManager(ocb.LogWrapper(env), batch_size=N)
    obs, st = mgr.batch_reset(ocr.PRNGKey(0))
    S, G = 8, 16
    keys = ocr.split_chain(ocr.PRNGKey(1), G)
    acts = [{"agent_0": torch.randint(0, 6, (N,), dtype=torch.int32, device=dev), "agent_1": torch.randint(0, 6, (N,), dtype=torch.int32, device=dev)} for _ in range(G)]
    outs_all = [{"obs_all": torch.empty((N, 2, 4, 5, 26), dtype=torch.uint8, device=dev)} for _ in range(S)]
    outs_two = [{"obs0": torch.empty((N, 4, 5, 26), dtype=torch.uint8, device=dev), "obs1": torch.empty((N, 4, 5, 26), dtype=torch.uint8, device=dev),
                 "reward": torch.empty(N, device=dev), "shaped0": torch.empty(N, device=dev), "shaped1": torch.empty(N, device=dev),
                 "done": torch.empty(N, dtype=torch.bool, device=dev)} for _ in range(S)]
    cat_out = [torch.empty((N, 2 * 520), dtype=torch.uint8, device=dev) for _ in range(S)]
    wenv = mgr._env

    def body_all(state):
        for i in range(G):
            o, state, r, d, info = mgr.batch_step(keys[i], state, acts[i], donate=True, out=outs_all[i % S])
        return state

    def body_cat(state):
        for i in range(G):
            o, state, r, d, info = wenv.step(ocr.LazySplit(keys[i], N), state, acts[i], donate=True, out=outs_two[i % S])
            torch.cat([o["agent_0"].reshape(N, -1), o["agent_1"].reshape(N, -1)], dim=-1, out=cat_out[i % S])
        return state

    res = {"envs": N}
    for name, body in (("kernel_written_all", body_all), ("two_arrays_then_torch_cat", body_cat)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            st = body(st); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                st = body(st)
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[name + "_us_per_step"] = round(e0.elapsed_time(e1) * 1e3 / (reps * G), 2)
    print(json.dumps(res), flush=True)
