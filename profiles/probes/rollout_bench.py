#!/usr/bin/env python
"""oc_rollout (T steps per launch) against the single-step graph, same box: us per step at several batch sizes (not product code).
usage: rollout_bench.py [--layout cramped_room] [--sizes 16,4096,65536,262144] [--T 16] [--packed]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import jaxovercooked_b200 as ocb
from jaxovercooked_b200 import random as ocr

ap = argparse.ArgumentParser()
ap.add_argument("--layout", default="cramped_room")
ap.add_argument("--sizes", default="16,256,4096,16384,65536,262144")
ap.add_argument("--T", default="16")
ap.add_argument("--packed", action="store_true")
ap.add_argument("--packed-in", action="store_true", help="uint8 actions only")
ap.add_argument("--packed-out", action="store_true", help="uint8 results only")
ap.add_argument("--random-reset", action="store_true")
ap.add_argument("--log", action="store_true")
ap.add_argument("--min-ms", type=float, default=150.0)
ap.add_argument("--single", action="store_true", help="also time the single-step graph of T launches")
args = ap.parse_args()
dev = torch.device("cuda")
lay = ocb.overcooked_layouts[args.layout]
env = ocb.make("overcooked", layout=lay, random_reset=args.random_reset)
B = env.info.algorithmic_bytes
PEAK = 6490.5e9
for T in [int(x) for x in args.T.split(",")]:
    for N in [int(x) for x in args.sizes.split(",")]:
        _, st = env.reset(ocr.split(ocr.PRNGKey(0), N))
        st.time.copy_(torch.randint(0, 400, (N,), dtype=torch.int32, device=dev))
        view = N * env.height * env.width * 26
        nbuf = max(2, min(8, int(300e6 // (2 * view * T)) + 1))      # rotate so that the obs written per cycle exceed L2
        obs = [torch.empty((T, 2, N, env.height, env.width, 26), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        nact = 4
        if args.packed or args.packed_in:
            acts = [(torch.randint(0, 6, (T, N), device=dev) | (torch.randint(0, 6, (T, N), device=dev) << 4)).to(torch.uint8) for _ in range(nact)]
        else:
            acts = [{"agent_0": torch.randint(0, 6, (T, N), dtype=torch.int32, device=dev),
                     "agent_1": torch.randint(0, 6, (T, N), dtype=torch.int32, device=dev)} for _ in range(nact)]
        if args.packed or args.packed_out:
            out = {"results8": torch.empty((T, N), dtype=torch.uint8, device=dev)}
        else:
            out = {"reward": torch.empty((T, N), device=dev), "shaped0": torch.empty((T, N), device=dev),
                   "shaped1": torch.empty((T, N), device=dev), "done": torch.empty((T, N), dtype=torch.bool, device=dev)}
        keys = ocr.split_chain(ocr.PRNGKey(1), T)
        log = None
        if args.log:
            log = {f: torch.zeros((N, 2), device=dev) for f in ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths")}
        stats = torch.zeros(8, dtype=torch.int64, device=dev) if args.log else None

        def run(i):
            env.rollout(keys, st, acts[i % nact], out=dict(out, obs=obs[i % nbuf]), log=log, stats=stats)
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters, ms = 8, 0.0
        while True:
            e0.record()
            for i in range(iters):
                run(i)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if ms >= args.min_ms or iters >= 1 << 16:
                break
            iters *= 2
        us = ms * 1e3 / (iters * T)
        line = {"kernel": "oc_rollout", "layout": args.layout, "N": N, "T": T, "us_per_step": round(us, 3), "env_steps_per_s": round(N / us * 1e6),
                "frac_1277": round(B * N / (us * 1e-6) / PEAK, 4), "packed": args.packed, "packed_in": args.packed_in, "packed_out": args.packed_out, "lib": os.path.basename(os.environ.get("OC_B200_LIB", "in-tree")), "log": args.log, "obs_bufs": nbuf}
        if args.single:
            slot = [{"obs0": obs[k % nbuf][t, 0], "obs1": obs[k % nbuf][t, 1], "reward": out.get("reward", torch.empty((T, N), device=dev))[t],
                     "shaped0": torch.empty(N, device=dev), "shaped1": torch.empty(N, device=dev), "done": torch.empty(N, dtype=torch.bool, device=dev)}
                    for k in range(nbuf) for t in range(T)]
            a32 = acts if not args.packed else [{"agent_0": (a & 15).to(torch.int32), "agent_1": (a >> 4).to(torch.int32)} for a in acts]
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                def body():
                    for k in range(nbuf):
                        for t in range(T):
                            env.step(ocr.LazySplit(keys[t], N), st, {"agent_0": a32[k % nact]["agent_0"][t], "agent_1": a32[k % nact]["agent_1"][t]},
                                     donate=True, out=slot[k * T + t], log=log, stats=stats)
                body(); torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=side):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            g.replay(); torch.cuda.synchronize()
            reps = 2
            while True:
                e0.record()
                for i in range(reps):
                    g.replay()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if ms >= args.min_ms or reps >= 1 << 14:
                    break
                reps *= 2
            us1 = ms * 1e3 / (reps * nbuf * T)
            line["single_step_us"] = round(us1, 3)
            line["single_frac_1277"] = round(B * N / (us1 * 1e-6) / PEAK, 4)
        print(json.dumps(line), flush=True)
        del obs, acts, st
        torch.cuda.empty_cache()
