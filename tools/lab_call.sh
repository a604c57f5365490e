#!/bin/bash
# one GPU call: the pipeline labs on libc3d_lab.so (make -C omni3d_b200/csrc lab [LABFLAGS=...])
cd "$(dirname "$0")/.."
TAG=${1:-lab}
mkdir -p gpurun_out
{
timeout 300 python tools/pipeline_lab.py fwd 0,2,16,18
timeout 300 python tools/pipeline_lab.py wgrad 0,1,2,3
} > gpurun_out/pipeline_lab_$TAG.txt 2>&1
cat gpurun_out/pipeline_lab_$TAG.txt
