#!/bin/bash
# IoU-only gpurun call: parity tests, sweep, ncu of the filter / clip kernels.
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python -m pytest tests/test_iou3d_gpu.py tests/test_abi.py -m gpu -q -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -30 > gpurun_out/pytest_iou_$TAG.log
timeout 300 python tools/bench_iou3d.py > gpurun_out/iou3d_sweep_$TAG.log 2>&1
cp gpurun_out/iou3d_sweep.json gpurun_out/iou3d_sweep_$TAG.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou3d_ -c 6 -o gpurun_out/ncu_iou_$TAG -f python tools/iou_ncu_target.py > gpurun_out/ncu_iou_$TAG.log 2>&1
tail -5 gpurun_out/pytest_iou_$TAG.log; cut -c1-160 gpurun_out/iou3d_sweep_$TAG.log
