#!/bin/bash
# gpu.sh [gpurun args...] -- CMD : gpurun with retries while the pod answers "busy" (exit code 3)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
