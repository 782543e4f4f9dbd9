#!/bin/bash
RB="timeout 200 python profiles/probes/rollout_bench.py --sizes 65536"
$RB --T 1,2,4,8,16,32,64,128 | cut -c1-150
echo "== no pre-pack"; OC_B200_PACK_ACTIONS=0 $RB --T 16,64,128 | cut -c1-150
echo "== packed"; $RB --T 16,64,128 --packed | cut -c1-150
echo "== 4096 envs"; timeout 100 python profiles/probes/rollout_bench.py --sizes 4096 --T 16,64,128 | cut -c1-150
