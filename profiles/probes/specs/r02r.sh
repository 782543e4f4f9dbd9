#!/bin/bash
echo "== rollout tests with the plain-store path forced on"
OC_B200_PLAIN_OBS=1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_c_harness.py -q --timeout 200 --timeout-method thread 2>&1 | tail -3
echo "== and forced off"
OC_B200_PLAIN_OBS=0 timeout 600 python -m pytest tests/test_gpu_rollout.py -q --timeout 200 --timeout-method thread 2>&1 | tail -3
RB="timeout 100 python profiles/probes/rollout_bench.py"
echo "== default (plain for small batches)"; $RB --sizes 16,256,1024,4096,8192,16384,65536 2>&1 | cut -c1-150
echo "== bulk forced"; OC_B200_PLAIN_OBS=0 $RB --sizes 16,256,1024,4096,8192 2>&1 | cut -c1-150
echo "== plain forced"; OC_B200_PLAIN_OBS=1 $RB --sizes 4096,8192,16384,65536 2>&1 | cut -c1-150
echo "== plain forced epw4/8 small"; OC_B200_PLAIN_OBS=1 OC_B200_RO_EPW=4 $RB --sizes 16,1024,4096,8192 2>&1 | cut -c1-150; OC_B200_PLAIN_OBS=1 OC_B200_RO_EPW=8 $RB --sizes 4096,8192,16384 2>&1 | cut -c1-150
