#!/bin/bash
# A/B of rollout variants
RB="timeout 60 python profiles/probes/rollout_bench.py --sizes 4096,65536"
echo "== in-tree (1 span buffer, 7 CTAs)"; $RB
echo "== ro_2buf"; OC_B200_LIB=$PWD/build/variants/ro_2buf.so $RB
echo "== ro8 (1 span buffer, 64 regs, 8 CTAs)"; OC_B200_LIB=$PWD/build/variants/ro8.so $RB
echo "== in-tree packed-in"; $RB --packed-in
echo "== in-tree packed-out"; $RB --packed-out
echo "== in-tree packed"; $RB --packed
echo "== ro8 packed"; OC_B200_LIB=$PWD/build/variants/ro8.so $RB --packed
echo "== in-tree 262k"; timeout 60 python profiles/probes/rollout_bench.py --sizes 262144
echo "== ro8 262k"; OC_B200_LIB=$PWD/build/variants/ro8.so timeout 60 python profiles/probes/rollout_bench.py --sizes 262144
