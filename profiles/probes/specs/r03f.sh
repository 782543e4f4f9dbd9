#!/bin/bash
RB="timeout 100 python profiles/probes/rollout_bench.py --sizes 65536 --T 32"
for L in easy_layout_padded bottleneck_large c_kitchen basic_cooperative asymm_advantages square_arena easy_layout; do
  for e in 16 8 4; do echo "== $L epw $e"; OC_B200_RO_EPW=$e $RB --layout $L --random-reset | cut -c60-150; done
done
