#!/bin/bash
# ncu evidence for the round-2 kernel (oc_env_kernel<MODE_ROLLOUT>): launch list + one --set full capture of a steady-state launch
# at 65,536 and at 262,144 envs.  Each profiled command has first exited 0 without ncu.
set -x
B="python bench.py --steps 64 --warmup 32 --no-cpu --e2e-steps 0 --min-ms 0 --no-stats-window"
timeout 120 $B > gpurun_out/r02_ncu_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oc_env_kernel --launch-skip 8 --launch-count 1 -f -o gpurun_out/r02_env_rollout_65536 $B > gpurun_out/r02_ncu_full.log 2>&1
B2="python bench.py --steps 64 --warmup 32 --no-cpu --e2e-steps 0 --min-ms 0 --no-stats-window --envs 262144 --rotate 2"
timeout 120 $B2 > gpurun_out/r02_ncu_plain_262144.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oc_env_kernel --launch-skip 8 --launch-count 1 -f -o gpurun_out/r02_env_rollout_262144 $B2 > gpurun_out/r02_ncu_full_262144.log 2>&1
ls -la gpurun_out/*.ncu-rep
