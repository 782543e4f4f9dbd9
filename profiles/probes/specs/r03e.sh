#!/bin/bash
RB="timeout 100 python profiles/probes/rollout_bench.py --sizes 65536 --T 64"
for L in forced_coord coord_ring smallest_kitchen; do
  echo "== $L default"; $RB --layout $L | cut -c1-150
  echo "== $L template in shared memory up to 8 KB"; OC_B200_TMPL_SMEM_MAX=8192 $RB --layout $L | cut -c1-150
done
echo "== basic_cooperative 7x9"; $RB --layout basic_cooperative | cut -c1-150; OC_B200_RO_EPW=8 $RB --layout basic_cooperative | cut -c1-150
