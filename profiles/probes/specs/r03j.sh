#!/bin/bash
B="timeout 200 python bench.py --no-cpu --e2e-steps 0 --no-stats-window --steps 2000 --warmup 64 --min-ms 200"
S='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"]/1e9,3), "G", round(d["ms_per_step"]*1e3,2), "us")'
for e in 8 16 4; do echo "== one launch per step, epw $e"; OC_B200_STEP_EPW=$e $B --rollout-steps 0 | python -c "$S"; done
for e in 8 16; do echo "== policy rollout step, env epw $e"; OC_B200_STEP_EPW=$e $B --policy | python -c "$S"; done
for e in 8 16; do echo "== VDN-like log-wrapper one-step, epw $e"; OC_B200_STEP_EPW=$e $B --rollout-steps 0 --log-wrapper | python -c "$S"; done
