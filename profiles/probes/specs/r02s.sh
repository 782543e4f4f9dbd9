#!/bin/bash
for n in 16 4096; do timeout 100 python profiles/probes/rollout_config4.py --envs $n --steps 128; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02s_launches_16.csv python profiles/probes/rollout_config4.py --envs 16 --steps 32 --rollouts 2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02s_launches_4096.csv python profiles/probes/rollout_config4.py --envs 4096 --steps 32 --rollouts 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for n in (16, 4096):
    rows = [r for r in csv.reader(open(f"gpurun_out/r02s_launches_{n}.csv")) if len(r) > 14 and r[0].isdigit()]
    agg = collections.defaultdict(list)
    for r in rows[60:]:
        agg[r[4][:60]].append(float(r[14]) / 1e3)
    for k, v in agg.items():
        print(n, f"{k:60s} n={len(v):4d} avg {sum(v)/len(v):7.2f} us  min {min(v):7.2f}")
PY
