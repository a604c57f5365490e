#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> [timeout] [script] [gpus] — re-submit until the pod has a slot (transient answers are not charged)
TAG=$1; TO=${2:-2400}; SCRIPT=${3:-tools/gpu_call.sh}; GPUS=${4:-1}
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TO -- "bash $SCRIPT $TAG $GPUS" > gpurun_out/call_$TAG.stdout 2>&1
  if ! grep -q "status=transient\|another call" gpurun_out/call_$TAG.stdout; then break; fi
  sleep 60
done
echo finished after $i tries
