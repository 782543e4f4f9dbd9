#!/bin/bash
# 2-GPU check of the bench (own arm + reference arm) exactly as the driver launches it
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m_bench_2gpu.json 2> gpurun_out/r02m_bench_2gpu.err; tail -c 2500 gpurun_out/r02m_bench_2gpu.json; tail -5 gpurun_out/r02m_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m_ref_2gpu.json 2> gpurun_out/r02m_ref_2gpu.err; tail -c 600 gpurun_out/r02m_ref_2gpu.json
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02m_bench_1gpu.json 2> gpurun_out/r02m_bench_1gpu.err; tail -c 300 gpurun_out/r02m_bench_1gpu.json
