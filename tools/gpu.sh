#!/bin/bash
# tools/gpu.sh <timeout_s> [--gpus N] -- '<command>' : gpurun with retries while the pod has no free box (exit 3)
t=$1; shift
for attempt in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu.sh] no box (attempt $attempt), retrying in 10 s" >&2
  sleep 10
done
exit 3
