"""One-off soak of oc_rollout's bit-exactness against the CPU oracle: every registered layout x both reset modes x both PRNG
modes x every tile shape (2 / 4 / 8 / 16 envs per warp, forced through OC_B200_RO_EPW), ragged batches, LogWrapper trackers and
statistics on odd runs, explicit or fused keys, separate or side-by-side observation layout.
usage: python profiles/probes/fuzz_rollout.py [--envs 700] [--steps 48]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import jaxovercooked_b200 as ocb                                            # noqa: E402
from jaxovercooked_b200 import random as ocr                                # noqa: E402
from jaxovercooked_b200.layouts import overcooked_layouts                   # noqa: E402
from oracle import oracle as orc                                            # noqa: E402
from tests import gpu_helpers as gh                                         # noqa: E402
from tests.helpers import LOG_FIELDS, assert_same                           # noqa: E402

ACTION_P = np.array([0.15, 0.15, 0.15, 0.15, 0.05, 0.35])


def one(lay, name, rr, part, epw, N, T, seed, log, fused_keys, interleaved):
    os.environ["OC_B200_RO_EPW"] = str(epw)
    impl = "threefry_partitionable" if part else "threefry_legacy"
    max_steps = 9 + seed % 11
    env = ocb.make("overcooked", layout=lay, random_reset=rr, max_steps=max_steps, prng_impl=impl)
    ref = orc.Oracle(lay, random_reset=rr, max_steps=max_steps, partitionable=part)
    rng = np.random.default_rng(seed)
    rk = rng.integers(0, 2**32, (N, 2), dtype=np.uint32)
    _, st = env.reset(gh.keys_to_dev(rk))
    _, rst = ref.reset(rk)
    off = rng.integers(0, max_steps, N).astype(np.int32)
    rst["time"][:] = off
    st.time.copy_(gh.to_dev(off))
    a = rng.choice(6, size=(T, N, 2), p=ACTION_P).astype(np.int32)
    acts = {"agent_0": gh.to_dev(a[:, :, 0]), "agent_1": gh.to_dev(a[:, :, 1])}
    if fused_keys:
        keys = ocr.split_chain(ocr.PRNGKey(seed), T, impl=impl)
        sk = np.stack([ocr.split(keys[t], N, impl=impl).cpu().numpy() for t in range(T)])
    else:
        sk = rng.integers(0, 2**32, (T, N, 2), dtype=np.uint32)
        keys = gh.keys_to_dev(sk)
    logd = rlog = stats = None
    if log:
        logd = {f: torch.zeros((N, 2), dtype=torch.float32, device="cuda") for f in LOG_FIELDS}
        rlog = ref.new_log(N)
        stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = {}
    if interleaved:
        out["obs_all"] = torch.zeros((T, N, 2, env.height, env.width, 26), dtype=torch.uint8, device="cuda")
    obs, _, rew, done, info = env.rollout(keys, st, acts, out=out, log=logd, stats=stats)
    o0, o1 = obs["agent_0"].cpu().numpy(), obs["agent_1"].cpu().numpy()
    rw, dn = rew["agent_0"].cpu().numpy(), done["__all__"].cpu().numpy()
    s0g, s1g = info["shaped_reward"]["agent_0"].cpu().numpy(), info["shaped_reward"]["agent_1"].cpu().numpy()
    what = f"{name} rr={rr} part={part} epw={epw}"
    for t in range(T):
        (r0, r1), rr_, s0, s1, dd = ref.step(sk[t], rst, a[t, :, 0], a[t, :, 1], log=rlog)
        assert_same(o0[t], r0, f"{what} t={t} obs0")
        assert_same(o1[t], r1, f"{what} t={t} obs1")
        assert_same(rw[t], rr_, f"{what} t={t} reward")
        assert_same(s0g[t], s0, f"{what} t={t} shaped0")
        assert_same(s1g[t], s1, f"{what} t={t} shaped1")
        assert_same(dn[t], dd, f"{what} t={t} done")
    gh.compare_state(st, rst, what)
    if log:
        for f in LOG_FIELDS:
            assert_same(logd[f].cpu().numpy(), rlog[f], f"{what} log {f}")
        assert int(stats[6]) == N * T
    return N * T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=700)
    ap.add_argument("--steps", type=int, default=48)
    a = ap.parse_args()
    t0, runs, steps = time.time(), 0, 0
    for i, (name, lay) in enumerate(sorted(overcooked_layouts.items())):
        for rr in (False, True):
            for k, epw in enumerate((2, 4, 8, 16)):
                n = i * 8 + rr * 4 + k
                steps += one(lay, name, rr, part=bool((i + k) % 2), epw=epw, N=a.envs + 13 * k + i, T=a.steps, seed=7000 + n,
                             log=bool(n % 2), fused_keys=bool((n // 2) % 2), interleaved=bool((n // 3) % 2))
                runs += 1
    print(f"fuzz ok: {runs} rollouts, {steps / 1e6:.1f} M env-steps compared bit-for-bit in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
