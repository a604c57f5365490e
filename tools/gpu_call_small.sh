#!/bin/bash
# short gpurun call: selected GPU tests + one bench line (no baselines).  usage: bash tools/gpu_call_small.sh <tag> [pytest args]
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python -m pytest ${TESTS:-tests/test_kernels_gpu.py tests/test_heads_gpu.py tests/test_select_gpu.py} -m gpu -q -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -15 > gpurun_out/pytest_small_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_small_$TAG.json 2> gpurun_out/bench_small_$TAG.err
tail -6 gpurun_out/pytest_small_$TAG.log; head -c 420 gpurun_out/bench_small_$TAG.json; echo; tail -3 gpurun_out/bench_small_$TAG.err
