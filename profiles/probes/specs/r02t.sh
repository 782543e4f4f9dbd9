#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/r02t_gpu_tests.log 2>&1; tail -6 gpurun_out/r02t_gpu_tests.log
for n in 16 256 4096; do timeout 100 python profiles/probes/rollout_config4.py --envs $n --steps 128; done
timeout 100 python profiles/probes/rollout_bench.py --single --sizes 16,256,4096,65536 | cut -c1-100,230-330
