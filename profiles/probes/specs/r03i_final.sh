#!/bin/bash
# what the driver runs at round end, on one box
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03i_ref.json 2> gpurun_out/r03i_ref.err ) 2>&1 | grep real
( time timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03i_bench.json 2> gpurun_out/r03i_bench.err ) 2>&1 | grep real
python profiles/probes/specs/summarize.py gpurun_out/r03i_bench.json
python - <<'PY'
import json
a=json.loads(open("gpurun_out/r03i_bench.json").read().strip().splitlines()[-1]); b=json.loads(open("gpurun_out/r03i_ref.json").read().strip().splitlines()[-1])
print("same config:", a["config"]==b["config"], "ratio e2e:", a["e2e"]["value"]/b["e2e"]["value"], "ratio value:", a["value"]/b["value"], "ref", b["value"], b["cpu_baseline"]["cores"])
print(a["roofline"]["frac"], a["clocks"])
PY
