#!/bin/bash
P="timeout 200 python bench.py --policy --no-cpu --steps 2000 --warmup 64 --min-ms 200"
S='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"]/1e9,3), "G", round(d["ms_per_step"]*1e3,2), "us", (d.get("roofline") or {}).get("avg_launch_us"))'
echo "== 32 warps, 1 branch"; $P | python -c "$S"
echo "== 32 warps, 2 branches"; $P --streams 2 | python -c "$S"
echo "== 16 warps, 1 branch"; OC_B200_LIB=$PWD/build/variants/pol16.so $P | python -c "$S"
echo "== 16 warps, 2 branches"; OC_B200_LIB=$PWD/build/variants/pol16.so $P --streams 2 | python -c "$S"
echo "== 16 warps, 4 branches"; OC_B200_LIB=$PWD/build/variants/pol16.so $P --streams 4 | python -c "$S"
