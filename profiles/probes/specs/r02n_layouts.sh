#!/bin/bash
# bench lines behind the per-layout / per-config numbers in DESIGN.md (1 GPU): one JSON line each into gpurun_out/r02n_layouts.jsonl
OUT=gpurun_out/r02n_layouts.jsonl; : > $OUT
B="timeout 200 python bench.py --no-cpu --e2e-steps 0 --no-stats-window --steps 2000 --warmup 64 --min-ms 200"
for L in smallest_kitchen forced_coord coord_ring easy_layout square_arena basic_cooperative c_kitchen counter_circuit asymm_advantages; do $B --layout $L >> $OUT 2>> gpurun_out/r02n.err; done
# BASELINE configs[2]: the five padded 5x9 CL layouts, one GPU's share of 262,144 envs over 8 GPUs, LogWrapper + statistics fused
for L in cramped_room asymm_advantages coord_ring forced_coord counter_circuit; do $B --layout $L --padded-cl --envs 32768 --log-wrapper >> $OUT 2>> gpurun_out/r02n.err; done
# BASELINE configs[4]: large / padded layouts, random_reset
$B --layout easy_layout_padded --random-reset >> $OUT 2>> gpurun_out/r02n.err
$B --layout bottleneck_large --random-reset >> $OUT 2>> gpurun_out/r02n.err
# BASELINE configs[3] shape: 4,096 envs with the LogWrapper; configs[0] shape: 16 envs
$B --envs 4096 --log-wrapper >> $OUT 2>> gpurun_out/r02n.err
$B --envs 16 --rotate 1 >> $OUT 2>> gpurun_out/r02n.err
# the one-launch-per-step path and the packed e2e for reference
$B --rollout-steps 0 >> $OUT 2>> gpurun_out/r02n.err
timeout 200 python bench.py --policy --no-cpu --steps 2000 --warmup 64 --min-ms 200 >> $OUT 2>> gpurun_out/r02n.err
python - <<'PY'
import json
for l in open("gpurun_out/r02n_layouts.jsonl"):
    if not l.startswith("{"): continue
    d = json.loads(l); r = d.get("roofline") or {}
    print(f"{d['config']['workload'][:70]:70s} lw={d['config'].get('log_wrapper')} {d['value']/1e9:7.3f} G  {d['ms_per_step']*1e3:7.2f} us  frac {r.get('frac', 0):.3f}  survey-bytes frac {r.get("frac_survey_bytes", 0):.3f}")
PY
tail -3 gpurun_out/r02n.err
