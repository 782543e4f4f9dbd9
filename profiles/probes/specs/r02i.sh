#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_c_harness.py "tests/test_gpu_parity.py::test_ct_rollout_manager_vdn_surface" -x -q --timeout 100 --timeout-method thread > gpurun_out/r02i_tests.log 2>&1; tail -15 gpurun_out/r02i_tests.log
RB="timeout 60 python profiles/probes/rollout_bench.py"
echo "== in-tree (action ring)"; $RB --sizes 16,4096,65536,262144
echo "== in-tree packed"; $RB --sizes 65536 --packed
echo "== epw2"; OC_B200_EPW=2 $RB --sizes 16,4096,65536
echo "== epw4"; OC_B200_EPW=4 $RB --sizes 16,4096,65536
echo "== epw16"; OC_B200_EPW=16 $RB --sizes 4096,65536
echo "== T=128"; $RB --sizes 4096,65536 --T 128
