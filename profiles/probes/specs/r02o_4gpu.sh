#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { n=$1; port=$2; out=$3; shift 3; timeout 200 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; grep -c '^{' gpurun_out/$out.json; tail -2 gpurun_out/$out.err | cut -c1-300; }
run 4 29602 r02q_bench_4gpu
run 4 29605 r02q_cfg4_10x10_4gpu --layout easy_layout_padded --random-reset
python profiles/probes/specs/summarize.py gpurun_out/r02q_*.json
