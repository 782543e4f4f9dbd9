#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02x_gpu_tests.log 2>&1; tail -5 gpurun_out/r02x_gpu_tests.log
bash profiles/probes/specs/r02_ncu.sh > gpurun_out/r02x_ncu.log 2>&1; tail -3 gpurun_out/r02x_ncu.log
