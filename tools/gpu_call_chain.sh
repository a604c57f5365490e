#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python -m pytest tests/test_grad_chain_gpu.py tests/test_model_gpu.py tests/test_kernels_gpu.py tests/test_trainer_split.py -m gpu -q -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -40 > gpurun_out/pytest_chain_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_chain_$TAG.json 2> gpurun_out/bench_chain_$TAG.err
C3D_NO_GRAD_CHAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_nochain_$TAG.json 2> gpurun_out/bench_nochain_$TAG.err
tail -25 gpurun_out/pytest_chain_$TAG.log; head -c 400 gpurun_out/bench_chain_$TAG.json; echo; head -c 400 gpurun_out/bench_nochain_$TAG.json; tail -5 gpurun_out/bench_chain_$TAG.err
