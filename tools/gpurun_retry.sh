#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> [timeout] — re-submit until the pod has a slot (transient answers are not charged)
TAG=$1; TO=${2:-2400}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "bash tools/gpu_call.sh $TAG" > gpurun_out/call_$TAG.stdout 2>&1
  if ! grep -q "status=transient" gpurun_out/call_$TAG.stdout; then break; fi
  sleep 90
done
echo finished after $i tries
