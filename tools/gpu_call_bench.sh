#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
head -c 600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
