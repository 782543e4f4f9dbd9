#!/bin/bash
RB="timeout 100 python profiles/probes/rollout_bench.py"
for e in 2 4 8 16; do echo "== epw$e"; OC_B200_EPW=$e $RB --sizes 16,256,1024,4096,8192,16384,32768,65536 2>&1 | cut -c1-150; done
echo "== evict_first epw8"; OC_B200_LIB=$PWD/build/variants/ro_ef.so $RB --sizes 65536,262144 | cut -c1-150
echo "== evict_first epw16"; OC_B200_EPW=16 OC_B200_LIB=$PWD/build/variants/ro_ef.so $RB --sizes 65536,262144 | cut -c1-150
echo "== epw16 262k"; OC_B200_EPW=16 $RB --sizes 262144 | cut -c1-150
echo "== epw16 packed"; OC_B200_EPW=16 $RB --sizes 65536 --packed | cut -c1-150
echo "== epw16 other layouts"; OC_B200_EPW=16 $RB --sizes 65536 --layout coord_ring | cut -c1-150; OC_B200_EPW=16 $RB --sizes 65536 --layout easy_layout_padded --random-reset | cut -c1-150
