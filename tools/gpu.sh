#!/bin/bash
# tools/gpu.sh [--gpus N] [--timeout S] -- '<command>'   : gpurun with retries while the pod is busy (exit code 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
