#!/bin/bash
# one GPU call: tests, the pipeline labs (libc3d_lab.so), a default bench line
cd "$(dirname "$0")/.."
TAG=${1:-lab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=240 --tb=short 2>&1 | grep -v "^    assert\|^E  " | tail -40 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
{
timeout 300 python tools/pipeline_lab.py fwd 0,1,2,4,8,16,17,18,20,24
timeout 300 python tools/pipeline_lab.py wgrad 0,1,2,3
C3D_WGRAD_NO_BIGBOX=1 timeout 300 python tools/pipeline_lab.py wgrad 0,1,2,4,8
C3D_WGRAD_N128=1 timeout 300 python tools/pipeline_lab.py wgrad 0,1,2
C3D_WGRAD_NO_BIGBOX=1 C3D_WGRAD_PIX64=1 timeout 300 python tools/pipeline_lab.py wgrad 0,1,2
} > gpurun_out/pipeline_lab_$TAG.txt 2>&1
cat gpurun_out/pipeline_lab_$TAG.txt
timeout 600 python bench.py --steps 20 --warmup 5 --skip-torch-baseline --skip-cpu-baseline --skip-iou > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
head -c 400 gpurun_out/bench_$TAG.json
