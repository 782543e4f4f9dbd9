#!/bin/bash
# weak-scaling line per N with the final code, launched exactly as the driver does (one box)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=$1
timeout 240 $TR --nproc-per-node $N --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r03c_bench_${N}gpu.json 2> gpurun_out/r03c_bench_${N}gpu.err
tail -2 gpurun_out/r03c_bench_${N}gpu.err | cut -c1-300
python profiles/probes/specs/summarize.py gpurun_out/r03c_bench_${N}gpu.json
