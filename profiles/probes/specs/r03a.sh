#!/bin/bash
( time timeout 400 python bench.py > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err ) 2>&1 | grep real; python profiles/probes/specs/summarize.py gpurun_out/r03a_bench.json; tail -3 gpurun_out/r03a_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r03a_bench.json').read().strip().splitlines()[-1]); print(d['roofline']); print(d['steps'], d['replays'], d['timed_window_ms'], d['e2e']['steps'], d['e2e']['host_calls'], d['gpu_launches'], d['config']['l2'])"
timeout 600 python -m pytest tests/test_bench_contract.py tests/test_gpu_rollout.py -q -m gpu -k "bench or host_rollout" 2>&1 | tail -3
( time timeout 300 python bench.py --steps 20 --warmup 5 --gpus 1 > gpurun_out/r03a_bench_driver.json 2> /dev/null ) 2>&1 | grep real; python profiles/probes/specs/summarize.py gpurun_out/r03a_bench_driver.json
