#!/bin/bash
timeout 900 python profiles/probes/fuzz_rollout.py 2>&1 | tail -3
echo "== compute-sanitizer memcheck (rollout: tile shapes, interleaved, packed; C harness)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_rollout.py -q -x -k "tile_shape or interleaved_rollout or packed or t1" --timeout 600 --timeout-method thread 2>&1 | tail -8
echo "== compute-sanitizer racecheck (one small rollout)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest "tests/test_gpu_rollout.py::test_rollout_every_tile_shape[cramped_room-8]" "tests/test_gpu_rollout.py::test_rollout_every_tile_shape[cramped_room-16]" -q -x --timeout 500 --timeout-method thread 2>&1 | tail -8
