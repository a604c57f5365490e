#!/bin/bash
# One gpurun call of round 2: GPU tests (no -x: every failure is listed), bench lines, per-shape conv table, ncu captures, launch list.
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short --timeout=240 --timeout-method=thread 2>&1 | grep -v '^E        +' > gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --config resnet34 --steps 10 --warmup 3 --skip-cpu-baseline --skip-iou > gpurun_out/bench_resnet34_$TAG.json 2> gpurun_out/bench_resnet34_$TAG.err
timeout 300 python bench.py --batch 16 --steps 10 --warmup 3 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_b16_$TAG.json 2> gpurun_out/bench_b16_$TAG.err
timeout 600 python tools/conv_shapes_bench.py > gpurun_out/conv_shapes_$TAG.log 2>&1
cp gpurun_out/conv_shapes.json gpurun_out/conv_shapes_$TAG.json 2>/dev/null
timeout 300 python tools/bench_iou3d.py > gpurun_out/iou3d_sweep_$TAG.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/profile_step_$TAG.log 2>&1
grep -E 'passed|failed|FAILED|Error' gpurun_out/pytest_$TAG.log | tail -30; head -c 600 gpurun_out/bench_$TAG.json; echo; head -c 300 gpurun_out/bench_resnet34_$TAG.json; echo; head -c 300 gpurun_out/bench_b16_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
