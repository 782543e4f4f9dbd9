#!/bin/bash
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py -q --timeout 200 --timeout-method thread -k "shard or logwrapper or packed" 2>&1 | tail -4
echo "== precision: in-tree (tanh.approx)"; timeout 120 python profiles/probes/policy_precision.py
echo "== precision: exact tanhf"; OC_B200_LIB=$PWD/build/variants/pol_exact.so timeout 120 python profiles/probes/policy_precision.py
echo "== policy bench exact tanhf"; OC_B200_LIB=$PWD/build/variants/pol_exact.so timeout 200 python bench.py --policy --no-cpu --steps 2000 --warmup 64 --min-ms 200 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline'])"
echo "== policy bench in-tree"; timeout 200 python bench.py --policy --no-cpu --steps 2000 --warmup 64 --min-ms 200 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline'])"
echo "== policy bench streams 2"; timeout 200 python bench.py --policy --streams 2 --no-cpu --steps 2000 --warmup 64 --min-ms 200 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
