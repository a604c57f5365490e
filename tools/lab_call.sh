#!/bin/bash
# one GPU call: tests, the two kernel labs, a default bench line
cd "$(dirname "$0")/.."
TAG=${1:-lab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=240 --tb=short 2>&1 | grep -v "^    assert\|^E  " | tail -40 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 900 bash tools/wgrad_lab.sh > /dev/null 2>&1
timeout 900 bash tools/conv_lab.sh > /dev/null 2>&1
timeout 600 python bench.py --skip-torch-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
head -c 600 gpurun_out/bench_$TAG.json
