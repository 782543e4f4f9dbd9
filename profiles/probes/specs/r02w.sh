#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_c_harness.py -q --timeout 200 --timeout-method thread 2>&1 | tail -3
RB="timeout 100 python profiles/probes/rollout_bench.py"
echo "== pre-pack (default)"; $RB --sizes 16384,65536,262144 | cut -c1-150
echo "== no pre-pack"; OC_B200_PACK_ACTIONS=0 $RB --sizes 16384,65536,262144 | cut -c1-150
echo "== packed input"; $RB --sizes 65536 --packed-in | cut -c1-150
echo "== other layouts"; $RB --sizes 65536 --layout coord_ring | cut -c1-150;  $RB --sizes 65536 --layout easy_layout_padded --random-reset | cut -c1-150
timeout 300 python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; python profiles/probes/specs/summarize.py gpurun_out/r02w_bench.json; python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1]); print(d['roofline']['frac'], d['roofline']['frac_kernel_bytes'], d['gpu_launches'], d['steps'])"
