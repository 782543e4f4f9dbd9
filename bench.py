#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched Overcooked step on B200, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   the reference's algorithm on the host cores (oracle)
    torchrun ... bench.py --gpus N ...                        one rank per GPU, env axis sharded, no per-step comms

A "step" is one pass of the hot path (split key -> step_env -> auto-reset -> both observations) over one batch
of `--envs` environments per GPU with synthetic uniform-random actions.  Workload = BASELINE.json configs[1]:
cramped_room, 65,536 envs per GPU.  One JSON line is printed by rank 0.

By default the steps run through `oc_rollout`, `--rollout-steps` (128 = the reference's scan length, baselines/IPPO_CL.py:42,577)
steps per launch with the environments resident on chip -- what the reference's `lax.scan` over `vmap(env.step)` does when the
actions are given; `--rollout-steps 0` launches one `oc_step` kernel per step (what a policy in the loop needs).

Timing hygiene: W >= 3 warm-up steps; `--rotate` independent env batches (state) are stepped round-robin and every launch
writes a fresh trajectory buffer far larger than the 126 MB L2 (one-launch-per-step mode: 8 observation slots round-robin);
the timed region is a replayed CUDA graph bracketed by barrier + synchronize, timed with CUDA events on the launching stream,
max over ranks, and never shorter than `--min-ms` (250 ms) whatever `--steps` says; the i.i.d. actions and the per-step subkeys
are re-drawn on the device inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "env-steps/sec (agent-pairs)"
UNIT = "env-steps/s"
OBS_SLOTS = 8
GRAPH_STEPS = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="cramped_room")
    ap.add_argument("--padded-cl", action="store_true", help="use the 5x9-padded CL version of --layout")
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU")
    ap.add_argument("--rotate", type=int, default=8, help="independent env batches stepped round-robin")
    ap.add_argument("--random-reset", action="store_true")
    ap.add_argument("--max-steps", type=int, default=400)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=4000,
                    help="steps of the end-to-end leg (host buffers through the C ABI); raised to 2000 unless --min-ms 0; 0 skips the leg")
    ap.add_argument("--min-ms", type=float, default=250.0,
                    help="floor of the device-timed window: the 16-step graph is replayed until the CUDA-event window is at least "
                         "this long, whatever --steps says (the line reports the steps actually timed); 0 = exactly --steps")
    ap.add_argument("--no-stats-window", action="store_true",
                    help="skip the LogWrapper-fused window whose int64[8] statistics are all-reduced over the ranks")
    ap.add_argument("--stats-steps", type=int, default=448)
    ap.add_argument("--rollout-steps", type=int, default=128,
                    help="steps per launch: oc_rollout keeps the environments on chip for this many steps (the callers' lax.scan "
                         "over env.step with the actions given; default = the reference's own scan length, num_steps = 128, "
                         "baselines/IPPO_CL.py:42,577); 0 = one oc_step launch per step in a replayed CUDA graph")
    ap.add_argument("--e2e-wire", default="packed", choices=["reference", "packed"],
                    help="host wire format of the e2e leg: the reference's dtypes (int32 actions, float32 results: 21 B per "
                         "env-step) or the packed one (2 B per env-step); the other one is reported next to it")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--log-wrapper", action="store_true", help="fuse the LogWrapper + statistics epilogue")
    ap.add_argument("--streams", type=int, default=1, choices=[1, 2, 4],
                    help="step the independent env batches on this many concurrent branches of the graph (default 1: every "
                         "launch waits for the previous one, as the steps of a single rollout do)")
    ap.add_argument("--emit-cells", action="store_true",
                    help="also write the compact observations (oc_step_extras.cells) every step, without a policy consuming them")
    ap.add_argument("--policy-input", default="cells", choices=["cells", "obs"],
                    help="with --policy: read the env kernel's compact observations (default) or scan the uint8 observations")
    ap.add_argument("--policy", action="store_true",
                    help="scope row f3: actions come from the actor-critic MLP (oc_policy_forward on the previous "
                         "observations) instead of the uniform stand-in; reports the whole rollout step")
    return ap.parse_args()


def get_layout(args):
    from jaxovercooked_b200.layouts import overcooked_layouts, cl_sequence_padded
    if args.padded_cl:
        return cl_sequence_padded()[args.layout]
    return overcooked_layouts[args.layout]


def workload_name(args, lay):
    return (f"{args.layout}{' (padded 5x9)' if args.padded_cl else ''} random-action env step, "
            f"{args.envs} envs per GPU, grid {lay['height']}x{lay['width']}, max_steps {args.max_steps}, "
            f"random_reset={bool(args.random_reset)}")


def workload_config(args, lay, world):
    """The `config` object: the workload only -- identical in our arm and in the reference arm (how each arm runs it is in `run`)."""
    return {"workload": workload_name(args, lay), "layout": args.layout, "padded_cl": bool(args.padded_cl),
            "grid": [int(lay["height"]), int(lay["width"])], "envs_per_gpu": args.envs, "global_envs": args.envs * world,
            "max_steps": args.max_steps, "random_reset": bool(args.random_reset), "actions": "i.i.d. uniform {0..5}",
            "log_wrapper": bool(args.log_wrapper), "policy": bool(args.policy)}


def alg_bytes(lay):
    """both obs written (52 hw) + interior of maze_map r/w (6 hw) + dynamic fields r/w (82)
    + key/actions in, reward/shaped/dones out (35)."""
    return 58 * int(lay["height"]) * int(lay["width"]) + 117


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(layout_name, envs=None, steps=None):
    """dram bytes per launch from the committed ncu --set full capture, if there is one for this workload
    (rollout launches are keyed by layout, envs per launch and steps per launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        key = layout_name if steps is None else f"{layout_name}_rollout_{envs}x{steps}"
        return t.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------- CPU arm

def host_threads():
    """Host threads this process may use: its CPU affinity mask (a launcher can narrow it), else the core count."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_rate(lay, args, n_envs, seconds, threads=None, steps=None):
    """env-steps/s of the oracle (OpenMP over envs) on the host cores."""
    from oracle import oracle as orc
    orc.set_threads(threads or host_threads())
    cores = orc.max_threads()
    o = orc.Oracle(lay, random_reset=args.random_reset, max_steps=args.max_steps)
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2**32, (n_envs, 2), dtype=np.uint32)
    _, st = o.reset(keys)
    out = o.alloc_out(n_envs)
    acts = rng.integers(0, 6, (16, n_envs, 2)).astype(np.int32)
    a0 = [np.ascontiguousarray(acts[i, :, 0]) for i in range(16)]
    a1 = [np.ascontiguousarray(acts[i, :, 1]) for i in range(16)]
    for i in range(3):
        o.step(keys, st, a0[i], a1[i], out=out)
    t0 = time.perf_counter()
    n = 0
    while True:
        o.step(keys, st, a0[n % 16], a1[n % 16], out=out)
        n += 1
        if steps is not None:
            if n >= steps:
                break
        elif time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return n_envs * n / dt, cores, n, dt


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on this box's host cores.

    1. If a real `jax` and the reference tree (baseline/_ref or /root/reference) are importable: the reference's
       jitted, vmapped step in a lax.scan on JAX_PLATFORMS=cpu (baseline/ref_jax.py; kind "reference-jax-cpu").
    2. Otherwise (this image: no jax wheel, no network): the C restatement oracle/oc_oracle.c with OpenMP (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lay = get_layout(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        time.sleep(5.0)                    # let the other (idle) ranks of the torchrun job import, see their rank and exit
    line = None
    try:
        from baseline import ref_jax
        if ref_jax.available():
            line = run_reference_jax(args, lay, ref_jax)
    except Exception as ex:                # a broken JAX install must not take the arm down: fall back to the port, say why
        print(f"reference arm: real-JAX path failed ({type(ex).__name__}: {ex}); timing the C port instead", file=sys.stderr)
        line = None
    if line is None:
        line = run_reference_port(args, lay)
    print(json.dumps(line), flush=True)


def reference_line(args, lay, value, steps, warmup, ms_per_step, n_envs, cores, kind, sample):
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, lay, max(1, args.gpus)), "run": {"sample_envs_per_step": n_envs},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def run_reference_jax(args, lay, ref_jax):
    jax = ref_jax.import_jax()
    ns = ref_jax.load_reference()
    cores = host_threads()
    steps = max(1, args.steps)
    # bounded sample: a pilot at 1,024 envs sizes the batch so that warm-up + 3 timed runs of `steps` steps take about a minute
    pilot, _, _ = ref_jax.time_random_rollout(jax, ns, lay, args.layout, n_envs=1024, n_steps=8, random_reset=args.random_reset,
                                              max_steps=args.max_steps, repeats=1)
    n_envs = int(min(args.envs, max(256, pilot * 15.0 / steps)))
    n_envs = max(256, (n_envs // 256) * 256)
    value, best, compile_s = ref_jax.time_random_rollout(jax, ns, lay, args.layout, n_envs=n_envs, n_steps=steps,
                                                         random_reset=args.random_reset, max_steps=args.max_steps, repeats=3)
    sample = (f"{n_envs} of {args.envs} envs per step x {steps} steps, best of 3 after one compile+warm-up run ({compile_s:.1f} s): "
              f"the reference's own jax.jit(jax.vmap(env.step)) in a lax.scan (jax {jax.__version__}, JAX_PLATFORMS=cpu, "
              f"{cores} host threads available)")
    return reference_line(args, lay, value, steps, steps, best / steps * 1e3, n_envs, cores, "reference-jax-cpu", sample)


def run_reference_port(args, lay):
    from oracle import oracle as orc
    orc.set_threads(host_threads())        # torchrun exports OMP_NUM_THREADS=1; the other ranks do no work
    cores = orc.max_threads()
    # bounded sample: size each step so that warmup + K steps finish in about a minute
    probe, _, _, _ = cpu_rate(lay, args, 8192, 1.0)
    budget = 60.0
    n_envs = int(min(args.envs, max(256, probe * budget / max(1, args.steps + args.warmup))))
    n_envs = max(256, (n_envs // 256) * 256)
    o = orc.Oracle(lay, random_reset=args.random_reset, max_steps=args.max_steps)
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2**32, (n_envs, 2), dtype=np.uint32)
    _, st = o.reset(keys)
    out = o.alloc_out(n_envs)
    acts = rng.integers(0, 6, (16, n_envs, 2)).astype(np.int32)
    a0 = [np.ascontiguousarray(acts[i, :, 0]) for i in range(16)]
    a1 = [np.ascontiguousarray(acts[i, :, 1]) for i in range(16)]
    for i in range(max(3, args.warmup)):
        o.step(keys, st, a0[i % 16], a1[i % 16], out=out)
    t0 = time.perf_counter()
    for i in range(args.steps):
        o.step(keys, st, a0[i % 16], a1[i % 16], out=out)
    dt = time.perf_counter() - t0
    value = n_envs * args.steps / dt
    sample = (f"{n_envs} of {args.envs} envs per step x {args.steps} steps, C restatement of the reference "
              f"(oracle/oc_oracle.c, OpenMP, {cores} threads); the reference's own JAX path cannot be installed here "
              f"(no jax wheel in the image; bench.py picks it up through baseline/ref_jax.py wherever `import jax` works)")
    return reference_line(args, lay, value, args.steps, max(3, args.warmup), dt / args.steps * 1e3, n_envs, cores, "port", sample)


# ------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:   # CUDA_VISIBLE_DEVICES may renumber: resolve the CUDA device through its UUID
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                     "applications_clocks": 0x2}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.01)

    def start(self):
        if self.nv is None:
            return
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        if self._th is not None:
            self._stop.set()
            self._th.join()
        if self.nv is not None and not self.samples:
            self._once()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- statistics window

STATS_SPLIT_NUM = 1 << 23     # per-env keys are split(key, 2**23)[global env index]: the same for any number of GPUs


def stats_window(args, env, dev, world, rank, N):
    """A LogWrapper-fused window whose exact integer statistics (oc_step_extras.stats) are reduced over the ranks: the one
    collective of this path (the reference's `tree_map(mean)` over rollout info, baselines/IPPO_CL.py:771-812).

    Every environment's trajectory here depends only on its GLOBAL index (reset key, per-step key and actions are functions of
    the global index), so rank 0's shard is the same whatever the world size, and the all-reduced vector must equal the sum
    of the per-rank vectors gathered separately."""
    import torch
    import torch.distributed as dist
    import jaxovercooked_b200 as ocb
    from jaxovercooked_b200 import random as ocr
    from jaxovercooked_b200.dist import reduce_stats, stats_to_metrics
    T = max(1, args.stats_steps)
    first = rank * N
    wenv = ocb.LogWrapper(env)
    ks = ocr.split(ocr.PRNGKey(777), 2)
    _, st = wenv.reset(ocr.split(ks[1], STATS_SPLIT_NUM, first=first, count=N))
    subs = ocr.split_chain(ks[0].clone(), T)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    h, w = env.height, env.width
    out = {"obs0": torch.empty((N, h, w, 26), dtype=torch.uint8, device=dev),
           "obs1": torch.empty((N, h, w, 26), dtype=torch.uint8, device=dev),
           "reward": torch.empty(N, dtype=torch.float32, device=dev), "shaped0": torch.empty(N, dtype=torch.float32, device=dev),
           "shaped1": torch.empty(N, dtype=torch.float32, device=dev), "done": torch.empty(N, dtype=torch.bool, device=dev)}
    gidx = torch.arange(first, first + N, dtype=torch.int64, device=dev)

    def actions(t):   # a fixed integer hash of (step, global env index): uniform-looking, world-size independent
        x = (gidx * 0x9E3779B1 + (t + 1) * 0x85EBCA6B) & 0xFFFFFFFF
        x = ((x ^ (x >> 15)) * 0x2C1B3C6D) & 0xFFFFFFFF
        x = ((x ^ (x >> 12)) * 0x297A2D39) & 0xFFFFFFFF
        x = x ^ (x >> 15)
        return {"agent_0": (x % 6).to(torch.int32), "agent_1": ((x >> 8) % 6).to(torch.int32)}

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for t in range(T):
        _, st, _, _, _ = wenv.step(ocr.LazySplit(subs[t], STATS_SPLIT_NUM, first, N), st, actions(t), donate=True, out=out,
                                   stats=stats)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    shard = stats.clone()
    res = {"steps": T, "envs_per_gpu": N, "global_envs": N * world, "launches": T,
           "eager_ms": ms, "keys_and_actions": "functions of the global env index (split(key, 2**23)[i]; integer hash of (t, i))",
           "stats_names": ["finished_episodes", "sum_reward", "sum_shaped_agent_0", "sum_shaped_agent_1",
                           "sum_returned_episode_returns", "sum_returned_episode_lengths", "env_steps", "reserved"]}
    if world > 1:
        gathered = [torch.zeros_like(shard) for _ in range(world)]
        dist.all_gather(gathered, shard)
        warm = torch.zeros(8, dtype=torch.int64, device=dev)
        reduce_stats(warm)
        torch.cuda.synchronize()
        dist.barrier()
        ev0.record()
        reduce_stats(stats)                      # the collective itself: one NCCL all-reduce of int64[8]
        ev1.record()
        torch.cuda.synchronize()
        first_us = ev0.elapsed_time(ev1) * 1e3
        reps = 50
        ev0.record()
        for _ in range(reps):
            reduce_stats(warm)
        ev1.record()
        torch.cuda.synchronize()
        per_rank = [g.cpu().tolist() for g in gathered]
        total = [sum(v[i] for v in per_rank) for i in range(8)]
        res.update({"stats_allreduced": stats.cpu().tolist(), "stats_shard0": per_rank[0],
                    "allreduce_equals_sum_of_shards": total == stats.cpu().tolist(),
                    "allreduce_us": first_us, "allreduce_us_avg50": ev0.elapsed_time(ev1) * 1e3 / reps,
                    "backend": "nccl int64[8] sum"})
    else:
        res.update({"stats_allreduced": shard.cpu().tolist(), "stats_shard0": shard.cpu().tolist(),
                    "allreduce_equals_sum_of_shards": True, "allreduce_us": None, "backend": "single rank: no collective"})
    res["metrics"] = stats_to_metrics(torch.tensor(res["stats_allreduced"]))
    return res


# ------------------------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import jaxovercooked_b200 as ocb
    from jaxovercooked_b200 import random as ocr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pinned_cores = None
    if world > 1:
        # one process per GPU on one host: give every rank its own slice of the host cores, so that eight Python loops and
        # their CUDA driver threads do not migrate over each other (the e2e leg is host-driven)
        try:
            cores = sorted(os.sched_getaffinity(0))
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            per = len(cores) // max(1, lw)
            if per >= 1:
                pinned_cores = cores[local * per:(local + 1) * per]
                os.sched_setaffinity(0, pinned_cores)
        except Exception:
            pinned_cores = None
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    lay = get_layout(args)
    N, R = args.envs, max(1, args.rotate)
    n_global = N * world
    first = rank * N
    env = ocb.make("overcooked", layout=lay, random_reset=args.random_reset, max_steps=args.max_steps)
    wenv = ocb.LogWrapper(env) if args.log_wrapper else env
    h, w = env.height, env.width
    B_alg = alg_bytes(lay)

    # R independent env batches (different seeds), sharded by GLOBAL env index: same trajectories for any world size
    rngs, states = [], []
    for b in range(R):
        rng = ocr.PRNGKey(b)
        ks = ocr.split(rng, 2)
        rngs.append(ks[0].clone())
        _, st = wenv.reset(ocr.split(ks[1], n_global, first=first, count=N))
        states.append(st)
    obs_all = [torch.empty((2 * N, h * w * 26), dtype=torch.uint8, device=dev) for _ in range(OBS_SLOTS)]   # agent-major rows
    outs = [{"obs0": obs_all[i][:N].view(N, h, w, 26), "obs1": obs_all[i][N:].view(N, h, w, 26),
             "reward": torch.empty(N, dtype=torch.float32, device=dev),
             "shaped0": torch.empty(N, dtype=torch.float32, device=dev),
             "shaped1": torch.empty(N, dtype=torch.float32, device=dev),
             "done": torch.empty(N, dtype=torch.bool, device=dev)} for i in range(OBS_SLOTS)]
    net = params = pol_out = None
    if args.emit_cells and not args.policy:
        for i in range(OBS_SLOTS):
            outs[i]["cells"] = env.empty_cells(N)
    if args.policy:
        from jaxovercooked_b200.policy import ActorCritic
        assert R == OBS_SLOTS, "--policy reads each env batch's previous observations: needs --rotate == OBS_SLOTS"
        net = ActorCritic(env.action_space().n, prng_impl="threefry_legacy" if env.prng_mode == 0 else "threefry_partitionable")
        params = net.init(1234, h * w * 26, device=dev)
        with torch.no_grad():
            params["params"]["Dense_2"]["kernel"].mul_(100.0)       # a head that is not near-uniform: actions follow the obs
        net.bind(params, env=env, shard=(first, n_global) if world > 1 else None)   # sharded draws: same noise as one device
        cells = [env.empty_cells(N) for _ in range(R)]
        for b in range(R):
            outs[b]["cells"] = cells[b]
        pol_out = [{"action": torch.empty(2 * N, dtype=torch.int32, device=dev), "log_prob": torch.empty(2 * N, dtype=torch.float32, device=dev),
                    "value": torch.empty(2 * N, dtype=torch.float32, device=dev)} for _ in range(R)]
        for b in range(R):
            o = env.get_obs(states[b].env_state if args.log_wrapper else states[b], cells=cells[b])
            obs_all[b][:N].copy_(o["agent_0"].reshape(N, -1))
            obs_all[b][N:].copy_(o["agent_1"].reshape(N, -1))
    # synthetic policy: i.i.d. uniform actions, re-drawn on the device for every block of GRAPH_STEPS steps
    # (one graph-safe torch RNG kernel per block; the env never sees the same action tensor twice)
    # Two halves: while the steps of one half run, the next half's actions and subkeys are produced on a side stream
    # (what a policy / the caller's key chain would do concurrently with the env).
    torch.manual_seed(1234 + rank)
    act_buf = torch.zeros((2, GRAPH_STEPS, 2, N), dtype=torch.int32, device=dev)
    acts = [[{"agent_0": act_buf[hf, i, 0], "agent_1": act_buf[hf, i, 1]} for i in range(GRAPH_STEPS)] for hf in range(2)]
    stats = torch.zeros(8, dtype=torch.int64, device=dev) if args.log_wrapper else None

    # per-step subkeys: the callers' `rng, _rng = split(rng)` chain, refreshed for GRAPH_STEPS steps at a time;
    # the per-env `split(_rng, num_envs)` is fused into the step kernel (LazySplit)
    KPS = 2 if args.policy else 1          # subkeys per step: with a policy, one for pi.sample and one for env.step (IPPO_MLP.py:429,447)
    subkeys = torch.zeros((2, KPS * GRAPH_STEPS, 2), dtype=torch.uint32, device=dev)
    from jaxovercooked_b200 import _lib as oclib
    import ctypes as C
    lib = oclib.load()
    chain_rng = rngs[0]

    def generate(hf):
        """Fresh i.i.d. actions and the next GRAPH_STEPS subkeys of the `rng, _rng = split(rng)` chain for half `hf`."""
        if not args.policy:
            act_buf[hf].random_(0, 6)
        oclib.check(lib.oc_prng_chain(env.prng_mode, chain_rng.data_ptr(), KPS * GRAPH_STEPS, subkeys[hf].data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "oc_prng_chain")

    RO = args.rollout_steps if (args.rollout_steps > 0 and not args.policy and not args.emit_cells and args.streams == 1) else 0
    if RO:
        # oc_rollout: ONE launch steps a batch RO times.  Each half has its own action block [2, RO, N] and subkeys [RO, 2]
        # (re-drawn on the side stream while the other half runs) and its own trajectory buffers: observations in the
        # learners' layout uint8[RO, 2, N, h, w, 26] (1.09 GB per half at 65,536 envs x 16 steps: far larger than L2)
        ro_act = torch.zeros((2, 2, RO, N), dtype=torch.int32, device=dev)
        ro_keys = torch.zeros((2, RO, 2), dtype=torch.uint32, device=dev)
        n_ro_bufs = 2 if RO * 2 * N * h * w * 26 > 150e6 else 4
        ro_out = [{"obs": torch.empty((RO, 2, N, h, w, 26), dtype=torch.uint8, device=dev),
                   "reward": torch.empty((RO, N), dtype=torch.float32, device=dev),
                   "shaped0": torch.empty((RO, N), dtype=torch.float32, device=dev),
                   "shaped1": torch.empty((RO, N), dtype=torch.float32, device=dev),
                   "done": torch.empty((RO, N), dtype=torch.bool, device=dev)} for _ in range(n_ro_bufs)]

        def generate(hf):   # noqa: F811 -- the rollout form of the generator above
            ro_act[hf].random_(0, 6)
            oclib.check(lib.oc_prng_chain(env.prng_mode, chain_rng.data_ptr(), RO, ro_keys[hf].data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "oc_prng_chain")

    step_no = [0]

    def one_rollout(hf):
        j = step_no[0]
        step_no[0] += 1
        b = j % R
        st_b = states[b]
        log = None
        if args.log_wrapper:
            log = {f: getattr(st_b, f) for f in ("episode_returns", "episode_lengths", "returned_episode_returns",
                                                 "returned_episode_lengths")}
            st_b = st_b.env_state
        env.rollout((ro_keys[hf], n_global, first), st_b, {"agent_0": ro_act[hf, 0], "agent_1": ro_act[hf, 1]},
                    out=ro_out[j % n_ro_bufs], log=log, stats=stats if args.log_wrapper else None)

    def one_step(hf, i):
        j = step_no[0]
        step_no[0] += 1
        b, slot = j % R, j % OBS_SLOTS
        key = ocr.LazySplit(subkeys[hf, KPS * i + KPS - 1], n_global, first, N)
        kw = dict(donate=True, out=outs[slot])
        if args.log_wrapper:
            kw["stats"] = stats
        a = acts[hf][i]
        if args.policy:        # pi, value = network.apply(params, obs_batch); action = pi.sample(seed); log_prob = pi.log_prob(action)
            if args.policy_input == "cells":    # the env kernel's compact observations: the 26-channel images are not re-read
                action, _, _ = net.act_cells(params, cells[b], subkeys[hf, KPS * i], out=pol_out[b])
            else:
                action, _, _ = net.act(params, obs_all[b], subkeys[hf, KPS * i], out=pol_out[b])
            a = {"agent_0": action[:N], "agent_1": action[N:]}
        _, states[b], _, _, _ = wenv.step(key, states[b], a, **kw)

    side = torch.cuda.Stream(dev)

    extra = [torch.cuda.Stream(dev) for _ in range(args.streams - 1)]

    def block(n_blocks):
        """One block = 2 x GRAPH_STEPS steps; each half's inputs are generated while the other half steps."""
        main = torch.cuda.current_stream(dev)
        for _ in range(n_blocks):
            for hf in range(2):
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    generate(1 - hf)
                for st_ in extra:
                    st_.wait_stream(main)
                if RO:
                    one_rollout(hf)
                for i in range(0 if RO else GRAPH_STEPS):
                    lane_ = i % args.streams        # --streams > 1: independent env batches advance concurrently
                    if lane_ == 0:
                        one_step(hf, i)
                    else:
                        with torch.cuda.stream(extra[lane_ - 1]):
                            one_step(hf, i)
                for st_ in extra:
                    main.wait_stream(st_)
                main.wait_stream(side)

    BLOCK = 2 * (RO if RO else GRAPH_STEPS)
    K = max(BLOCK, (args.steps // BLOCK) * BLOCK)
    Wm = max(3, args.warmup)
    n_warm_blocks = max(2, (Wm + BLOCK - 1) // BLOCK)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        generate(0)
        block(1)                      # eager once (sets function attributes, warms allocator)
        torch.cuda.synchronize()
        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                block(1)
        run_blocks = (lambda n: [graph.replay() for _ in range(n)]) if graph is not None else block
        evw0, evw1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run_blocks(1)
        evw0.record(stream)
        run_blocks(n_warm_blocks)
        evw1.record(stream)
        torch.cuda.synchronize()
        # time floor: however few steps were asked for, the timed window is at least --min-ms long (same count on every rank)
        n_blocks = K // BLOCK
        if args.min_ms > 0:
            est = torch.tensor([evw0.elapsed_time(evw1) / n_warm_blocks], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(est, op=dist.ReduceOp.MIN)
            # (the warm-up estimate runs a little slow -- first replays -- so aim 20 % above the floor)
            n_blocks = max(n_blocks, int(np.ceil(1.2 * args.min_ms / max(float(est.item()), 1e-3))))
        K = n_blocks * BLOCK
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(stream)
        run_blocks(K // BLOCK)
        ev1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop()
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        if stats is not None and world > 1:
            dist.all_reduce(stats)     # the rollout statistics: the only collective on this path
    policy_us = None
    if args.policy and args.policy_input == "cells":
        # the policy kernel alone, on the compact observations the timed region left behind (its own roofline line below)
        evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            for b in range(R):
                net.act_cells(params, cells[b], subkeys[0, 0], out=pol_out[b])
            evp0.record(stream)
            reps_p = 25
            for _ in range(reps_p):
                for b in range(R):
                    net.act_cells(params, cells[b], subkeys[0, 0], out=pol_out[b])
            evp1.record(stream)
        torch.cuda.synchronize()
        policy_us = evp0.elapsed_time(evp1) * 1e3 / (reps_p * R)
    stats_win = None
    if not args.no_stats_window and not args.policy:
        stats_win = stats_window(args, env, dev, world, rank, N)

    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = n_global * K / (ms * 1e-3)
    # rollout mode: 2 oc_rollout + 2 key-chain kernels (+ 2 action pre-pack kernels for large int32 batches, oc_rollout_io)
    launches_per_block = ((2 if args.policy else 1) * BLOCK + 2) if not RO else (4 + (2 if (N >= 32768 and h * w <= 64) else 0))   # our kernels: 16 steps (+ 16 policy forwards) + 2 key-chain kernels (the 2 RNG kernels are torch's)
    gpu_launches = (K // BLOCK) * launches_per_block + (stats_win["launches"] if stats_win else 0)

    # ---- e2e: the C-ABI call with HOST buffers (oc_step_host through env.host_stepper): every step copies that
    # step's actions from pinned host memory to the device and the step's rewards / shaped rewards / dones back to
    # pinned host memory.  The R env batches run on R streams, so one batch's copies overlap another's kernel; the
    # host waits for a batch's previous results before it reuses that batch's buffers.
    e2e = None
    if RO and not args.log_wrapper and args.e2e_steps > 0:
        # ---- e2e, rollout form: oc_rollout_host through env.HostRollout.  Per RO steps of a batch: ONE host-to-device copy of
        # the RO steps' actions from pinned host memory, one launch, ONE device-to-host copy of the RO steps' rewards / shaped
        # rewards / dones into pinned host memory; the R env batches run on R streams.  Measured for both wire formats.
        Ke = max(2 * R * RO, args.e2e_steps if args.min_ms <= 0 else max(args.e2e_steps, 2000))
        if args.min_ms > 0:          # the same time floor as the device-timed window
            Ke = max(Ke, int(args.min_ms * 1e-3 / max(ms * 1e-3 / K, 1e-9)))
        # every concurrently stepping batch writes its OWN trajectory buffer: as many batches as fit next to the two buffers of
        # the device-timed leg (all 8 for the default workload; fewer for layouts whose 128-step trajectories are tens of GB)
        torch.cuda.synchronize()
        buf_bytes = RO * 2 * N * h * w * 26
        free_b = torch.cuda.mem_get_info(dev)[0]
        Re = int(max(2, min(R, 2 + (0.85 * free_b) // buf_bytes)))
        e_obs = [ro_out[i]["obs"] for i in range(min(2, n_ro_bufs))]
        e_obs += [torch.empty_like(e_obs[0]) for _ in range(Re - len(e_obs))]
        n_ro = (Ke + RO - 1) // RO
        n_ro = ((n_ro + Re - 1) // Re) * Re
        Ke = n_ro * RO
        e_keys = ocr.split_chain(rngs[1 % R], (n_ro + 2 * Re) * RO).view(n_ro + 2 * Re, RO, 2)
        res_e2e = {}
        for wire in ("reference", "packed"):
            packed = wire == "packed"
            hrs = [ocb.HostRollout(env, states[b], RO, n_global, first, packed=packed, obs=e_obs[b]) for b in range(Re)]
            if packed:
                h_pool = [(torch.randint(0, 6, (RO, N)) | (torch.randint(0, 6, (RO, N)) << 4)).to(torch.uint8).pin_memory() for _ in range(4)]
            else:
                h_pool = [torch.randint(0, 6, (2, RO, N), dtype=torch.int32).pin_memory() for _ in range(4)]
            torch.cuda.synchronize()

            def e2e_run(i0, n):
                for i in range(i0, i0 + n):
                    hr = hrs[i % Re]
                    hr.wait()                                    # this batch's previous results are on the host
                    hr.run(e_keys[i], h_pool[i % 4])

            e2e_run(0, 2 * Re)
            for hr in hrs:
                hr.wait()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_run(2 * Re, n_ro)
            for hr in hrs:
                hr.wait()
            dt = time.perf_counter() - t0
            tm = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            res_e2e[wire] = n_global * Ke / float(tm.item())
            del hrs
        wire = args.e2e_wire
        other = "packed" if wire == "reference" else "reference"
        bi, bo = (N, N) if wire == "packed" else (8 * N, 13 * N)
        e2e = {"value": res_e2e[wire], "unit": UNIT, "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo, "steps": Ke,
               "wire": wire, f"value_{other}_wire": res_e2e[other],
               f"bytes_per_env_step_{other}_wire": 2 if other == "packed" else 21,
               "host_calls": n_ro, "steps_per_host_call": RO, "pinned_cores": len(pinned_cores) if pinned_cores else None,
               "device_note": "the packed-wire launch reads 1-byte actions and writes 1-byte results, which is ~10 % faster on the "
                              "device than the int32 / float32 arrays of the device-timed `value`, so e2e (packed) can exceed `value`",
               "note": f"oc_rollout_host (C ABI, host buffers) via env.HostRollout: per {RO} steps of a batch one copy of the "
                       "actions from pinned host memory in, one launch, one copy of rewards+shaped+dones to pinned host memory "
                       "out (those bytes are counted per step); observations stay on the device for the learner, as in the "
                       f"reference; {Re} env batches pipelined on {Re} streams, each with its own trajectory buffer"}
    elif not args.log_wrapper and not args.policy and args.e2e_steps > 0:
        Ke = max(2 * R, args.e2e_steps if args.min_ms <= 0 else max(args.e2e_steps, 2000))
        torch.cuda.synchronize()
        steppers = [env.host_stepper(states[b], n_global, first) for b in range(R)]
        h_pool = [torch.randint(0, 6, (2, N), dtype=torch.int32).pin_memory() for _ in range(16)]
        e_subs = ocr.split_chain(rngs[1 % R], Ke + 2 * R)
        torch.cuda.synchronize()

        def e2e_run(i0, n):
            for i in range(i0, i0 + n):
                st_ = steppers[i % R]
                st_.wait()                                   # results of this batch's previous step are on the host
                st_.step(e_subs[i], h_pool[i % 16])

        e2e_run(0, 2 * R)
        for s_ in steppers:
            s_.wait()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_run(2 * R, Ke)
        for s_ in steppers:
            s_.wait()
        dt = time.perf_counter() - t0
        tm = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e = {"value": n_global * Ke / float(tm.item()), "unit": UNIT, "h2d_bytes_per_step": 8 * N,
               "d2h_bytes_per_step": 13 * N, "steps": Ke,
               "note": "oc_step_host (C ABI, host buffers) via env.host_stepper: actions from pinned host memory in, "
                       "rewards+shaped+dones to pinned host memory out, every step; observations stay on the device for "
                       f"the learner, as in the reference; {R} env batches pipelined on {R} streams"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    step_s = ms * 1e-3 / K
    achieved = B_alg * N / step_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic(args.layout), "kernel": "oc_env_kernel<MODE_STEP>",
                "algorithmic_bytes_per_env_step": B_alg, "envs_per_launch": N, "avg_launch_us": step_s * 1e6,
                "peak_source": peak_src}
    if RO:
        # one launch = N envs x RO steps.  SURVEY.md 8(d)'s per-env-step figure (1277 B for cramped_room) is per env.step CALL:
        # state in + state out + both observations + keys/actions/results.  The rollout kernel moves the state once per RO steps,
        # so the bytes IT must move per env-step are fewer: 52hw (observations) + 21 (actions in, results out) + (6hw + 82)/RO
        # (state).  `achieved` / `frac` use that smaller, honest figure -- with SURVEY's figure the "fraction" would count bytes
        # that are never moved (and exceeds 1 for long rollouts); it is reported next to it as *_survey_bytes.
        hw_ = h * w
        b_kernel = 52 * hw_ + 21 + (6 * hw_ + 82) / RO
        ach_k = b_kernel * N / step_s / 1e9
        roofline.update({"kernel": "oc_env_kernel<MODE_ROLLOUT>", "env_steps_per_launch": N * RO, "steps_per_launch": RO,
                         "avg_launch_us": step_s * 1e6 * RO, "traffic": profiled_traffic(args.layout, N, RO),
                         "algorithmic_bytes_per_env_step": b_kernel, "achieved": ach_k, "frac": ach_k / peak,
                         "survey_8d_bytes_per_env_step": B_alg, "achieved_survey_bytes": achieved,
                         "frac_survey_bytes": achieved / peak})
    cpu = None
    if world == 1 and not args.no_cpu and not args.policy:
        n_cpu = N          # the same batch the reference arm (`--impl reference`) steps
        v, cores, n_steps, dt = cpu_rate(lay, args, n_cpu, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} envs x {n_steps} steps in {dt:.1f} s, oracle/oc_oracle.c (C restatement, OpenMP); "
                         "the reference's JAX CPU path is not installable in this image"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": (n_warm_blocks + 2) * BLOCK,
            "requested_steps": args.steps, "replays": K // BLOCK, "timed_window_ms": ms,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": workload_config(args, lay, world),
            "run": {"l2": (f"inputs larger than L2: {R} independent env batches + {OBS_SLOTS} observation slots "
                              f"round-robin (state {R * N * (env.info.map_bytes_per_env + 41) / 1e6:.0f} MB + obs "
                              f"{OBS_SLOTS * 2 * N * h * w * 26 / 1e6:.0f} MB per cycle vs 126 MB L2)") if not RO else
                             (f"outputs larger than L2: every launch writes a fresh uint8[{RO}, 2, N, h, w, 26] trajectory buffer "
                              f"({RO * 2 * N * h * w * 26 / 1e6:.0f} MB; {n_ro_bufs} of them round-robin) and steps one of {R} "
                              f"independent env batches (state {R * N * (env.info.map_bytes_per_env + 41) / 1e6:.0f} MB per cycle) vs 126 MB L2"),
                       "launch": (f"CUDA graph of 2 oc_rollout launches x {RO} steps replayed (the next {RO} steps' actions/subkeys "
                                  "generated on a side branch)" if RO else
                                  "CUDA graph of 16 steps replayed (next 8 steps' actions/subkeys generated on a side branch)")
                                 if graph is not None else "eager launches",
                       "steps_per_launch": RO if RO else 1,
                       "keys": "per-step subkeys from the split chain; per-env split fused in the kernel",
                       "actions": f"i.i.d. uniform {{0..5}}, re-drawn on the device every {RO if RO else 8} steps inside the timed region",
                       "concurrent_branches": args.streams, "log_wrapper": bool(args.log_wrapper), "emit_cells": bool(args.emit_cells or args.policy), "parallelism": f"env-sharded x{world}, no per-step comms"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks}
    if args.policy:
        line["metric"] = METRIC + ", rollout step = actor-critic forward + sample + env step"
        # two kernels per step.  The env kernel's HBM roofline is the default run's; the policy kernel is bound by instruction
        # issue, not by HBM or the tensor pipe: warp-instructions per row (ncu, profiles/r01_ncu_policy_kernel.txt)
        # x rows / kernel time against 4 issue slots per SM-cycle
        line["roofline"] = None
        if policy_us is not None:
            WI_ROW = 317.0
            f_sm = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
            slots = 148 * 4 * f_sm
            ach = WI_ROW * 2 * N / (policy_us * 1e-6)
            line["roofline"] = {"bound": "issue", "kernel": "oc_policy_kernel<tanh, compact source>", "achieved": ach / 1e12,
                                "peak": slots / 1e12, "unit": "T warp-instr/s", "frac": ach / slots, "traffic": None,
                                "warp_instructions_per_row": WI_ROW, "rows_per_launch": 2 * N, "avg_launch_us": policy_us,
                                "note": "issue-slot bound: 148 SMs x 4 schedulers x SM clock; DRAM < 1 % and tensor pipe < 10 % busy "
                                        "for this kernel (profiles/r01_ncu_policy_kernel.txt)"}
        line["run"]["actions"] = ("sampled on the device from the actor-critic MLP (architectures/mlp.py, random orthogonal "
                                     "init, fp16 tensor-core layer 2) evaluated on the previous step's observations")
        line["run"]["policy_rows_per_step"] = 2 * N
        line["run"]["policy_input"] = args.policy_input
    if stats is not None:
        line["stats"] = stats.cpu().tolist()
    if stats_win is not None:
        stats_win.pop("launches", None)
        line["stats_window"] = stats_win
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as ge
        if not os.path.exists(ge.LIB):
            ge.build()
        run_ours(args)


if __name__ == "__main__":
    main()
