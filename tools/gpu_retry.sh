#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   (retries while the pod answers "busy"; nothing is charged for those)
# extra gpurun flags (e.g. --gpus 2) via GPURUN_FLAGS
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gave up: pod busy"; exit 3
