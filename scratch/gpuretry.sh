#!/bin/bash
# retries a gpurun call until the pod has a free slot: scratch/gpuretry.sh LOG TIMEOUT 'command'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then break; fi
  sleep 120
done
