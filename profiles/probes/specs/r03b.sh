#!/bin/bash
bash profiles/probes/specs/r02n_layouts.sh 2>&1 | tail -24
# launch list + DRAM bytes of the default bench (128 steps per launch): one pass, no replay
B="python bench.py --steps 256 --warmup 256 --no-cpu --e2e-steps 0 --min-ms 0 --no-stats-window"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r03_launches_T128.csv $B > gpurun_out/r03b_ncu.log 2>&1
grep "oc_env_kernel<(int)3" gpurun_out/r03_launches_T128.csv | tail -6 | cut -c100-400
