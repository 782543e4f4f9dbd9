#!/bin/bash
RB="timeout 100 python profiles/probes/rollout_bench.py --sizes 65536 --T 32"
for L in cramped_room smallest_kitchen forced_coord easy_layout square_arena asymm_advantages basic_cooperative c_kitchen easy_layout_padded bottleneck_large counter_circuit; do echo "== $L"; $RB --layout $L --random-reset | cut -c60-150; done
timeout 300 python -m pytest tests/test_gpu_rollout.py -q -x 2>&1 | tail -2
