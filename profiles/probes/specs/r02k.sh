#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method thread > gpurun_out/r02k_gpu_tests.log 2>&1; tail -8 gpurun_out/r02k_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 2500 gpurun_out/r02k_bench.json; tail -3 gpurun_out/r02k_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; tail -c 1500 gpurun_out/r02k_bench_ref.json
bash profiles/probes/specs/r02_ncu.sh > gpurun_out/r02k_ncu.log 2>&1; tail -5 gpurun_out/r02k_ncu.log
