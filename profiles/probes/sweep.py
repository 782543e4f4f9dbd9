#!/usr/bin/env python
"""A/B sweeps of kernel variants on one GPU: runs bench.py once per (library, env overrides, extra args) and prints a table.
usage: sweep.py OUT.txt  < spec lines:  label | lib-or-'-' | ENV=V,ENV=V | extra bench args"""
import json, os, subprocess, sys, time
out = open(sys.argv[1], "a")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for line in sys.stdin:
    line = line.strip()
    if not line or line.startswith("#"):
        continue
    label, lib, envs, extra = [x.strip() for x in (line.split("|") + ["", "", ""])[:4]]
    env = dict(os.environ)
    if lib and lib != "-":
        env["OC_B200_LIB"] = os.path.join(ROOT, "build", "variants", lib + ".so")
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=")
        env[k] = v
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--e2e-steps", "0", "--no-stats-window", "--min-ms", "150",
           "--steps", "16", "--warmup", "32"] + extra.split()
    t0 = time.time()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        msg = (f"{label:28s} {j['value']/1e9:7.3f} G  {j['ms_per_step']*1e3:7.2f} us  frac {j['roofline']['frac'] if j.get('roofline') else 0:.3f}  "
               f"clk {j['clocks']['sm_mhz']}  steps {j['steps']}  ({time.time()-t0:.0f}s)")
    except Exception as ex:
        msg = f"{label:28s} FAILED rc={r.returncode} {ex} :: {r.stderr[-300:]!r}"
    print(msg, flush=True)
    out.write(msg + "\n"); out.flush()
