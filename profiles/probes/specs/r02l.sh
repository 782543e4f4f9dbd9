#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/r02l_gpu_tests.log 2>&1; tail -12 gpurun_out/r02l_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -c 1800 gpurun_out/r02l_bench.json; tail -3 gpurun_out/r02l_bench.err
