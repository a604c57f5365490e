#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python -m pytest tests/test_grad_chain_gpu.py tests/test_reference_call_patterns_gpu.py -m gpu -q -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -40 > gpurun_out/pytest_t_$TAG.log
tail -25 gpurun_out/pytest_t_$TAG.log
