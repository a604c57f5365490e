#!/bin/bash
# One gpurun call of round 2: GPU tests (no -x: every failure is listed), bench line, per-shape conv table, launch list.
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short --timeout=240 --timeout-method=thread 2>&1 | grep -v '^E        +' > gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 600 python tools/conv_shapes_bench.py > gpurun_out/conv_shapes_$TAG.log 2>&1
cp gpurun_out/conv_shapes.json gpurun_out/conv_shapes_$TAG.json 2>/dev/null
timeout 300 python tools/linear_bench.py > gpurun_out/linear_bench_$TAG.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > gpurun_out/profile_step_$TAG.log 2>&1
grep -E 'passed|failed|FAILED|Error' gpurun_out/pytest_$TAG.log | tail -30; head -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
