#!/bin/bash
RB="timeout 100 python profiles/probes/rollout_bench.py --sizes 65536,262144"
echo "== 4 warps per CTA (default)"; $RB | cut -c1-150
for w in 3 2 1; do echo "== $w warps per CTA"; OC_B200_RO_WARPS=$w $RB | cut -c1-150; done
echo "== 2 warps, epw 8"; OC_B200_RO_WARPS=2 OC_B200_RO_EPW=8 $RB | cut -c1-150
echo "== T=64, default"; timeout 100 python profiles/probes/rollout_bench.py --sizes 65536 --T 64 | cut -c1-150
