#!/bin/bash
( time timeout 400 python bench.py > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err ) 2>&1 | grep real; python profiles/probes/specs/summarize.py gpurun_out/r03d_bench.json; tail -3 gpurun_out/r03d_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r03d_bench.json').read().strip().splitlines()[-1]); print(d['e2e']['note'][-120:], d['e2e']['steps'], d['e2e']['host_calls'])"
timeout 300 python bench.py --layout easy_layout_padded --random-reset --no-cpu > gpurun_out/r03d_bench_10x10.json 2> gpurun_out/r03d_bench_10x10.err; python profiles/probes/specs/summarize.py gpurun_out/r03d_bench_10x10.json; tail -2 gpurun_out/r03d_bench_10x10.err
timeout 600 python -m pytest tests/test_bench_contract.py -q -m gpu 2>&1 | tail -3
