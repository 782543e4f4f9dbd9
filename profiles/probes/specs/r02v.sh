#!/bin/bash
RB="timeout 100 python profiles/probes/rollout_bench.py"
echo "== in-tree"; $RB --sizes 65536,262144 | cut -c1-150
echo "== L2 prefetch of the actions"; OC_B200_LIB=$PWD/build/variants/ro_pf1.so $RB --sizes 65536,262144 | cut -c1-150
echo "== L2 prefetch evict_last"; OC_B200_LIB=$PWD/build/variants/ro_pf2.so $RB --sizes 65536,262144 | cut -c1-150
echo "== packed"; $RB --sizes 65536 --packed | cut -c1-150
echo "== in-tree again"; $RB --sizes 65536 | cut -c1-150
