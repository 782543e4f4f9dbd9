#!/bin/bash
# multi-GPU evidence on one 8 x B200 box, launched exactly as the driver does (2 runs: budget)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { n=$1; port=$2; out=$3; shift 3; timeout 200 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; grep -c '^{' gpurun_out/$out.json; tail -2 gpurun_out/$out.err | cut -c1-300; }
run 8 29601 r02o_bench_8gpu
run 8 29603 r02o_cfg2_coord_ring_8gpu --layout coord_ring --padded-cl --envs 32768 --log-wrapper
nvidia-smi topo -m > gpurun_out/r02o_topo.txt 2>&1
python profiles/probes/specs/summarize.py gpurun_out/r02o_*.json
