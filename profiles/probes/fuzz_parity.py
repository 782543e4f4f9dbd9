"""One-off soak of the bit-exact parity check (same comparison as tests/test_gpu_parity.py::_rollout_compare) at larger
batches, longer horizons and more seeds than the test-suite affords: every registered layout x both reset modes x both PRNG
modes, LogWrapper fused on odd runs.  usage: python profiles/probes/fuzz_parity.py [--envs 1500] [--steps 150] [--seeds 2]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from jaxovercooked_b200.layouts import overcooked_layouts                   # noqa: E402
from tests.test_gpu_parity import _rollout_compare                          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1500)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--seeds", type=int, default=2)
    a = ap.parse_args()
    t0, runs, steps = time.time(), 0, 0
    for i, (name, lay) in enumerate(sorted(overcooked_layouts.items())):
        for s in range(a.seeds):
            for rr in (False, True):
                part = bool((i + s) % 2)
                _rollout_compare(lay, random_reset=rr, max_steps=23 + 7 * s, N=a.envs + 37 * s, T=a.steps, seed=1000 * i + 10 * s + rr,
                                 part=part, every=5, log=bool((i + s + rr) % 2))
                runs += 1
                steps += (a.envs + 37 * s) * a.steps
    print(f"fuzz ok: {runs} rollouts, {steps / 1e6:.1f} M env-steps compared bit-for-bit in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
