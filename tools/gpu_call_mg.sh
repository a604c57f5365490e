#!/bin/bash
# Multi-GPU call (gpurun --gpus N): weak scaling at 32 img/GPU and BASELINE configs[2]'s 16 img/GPU, two-stage backward + overlapped
# early-bucket all-reduce.  usage: bash tools/gpu_call_mg.sh <tag> <ngpus>
mkdir -p gpurun_out
TAG=${1:-mg}; N=${2:-2}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_n${N}_b32_$TAG.json 2> gpurun_out/bench_n${N}_b32_$TAG.err
run --batch 16 --steps 20 --warmup 5 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_n${N}_b16_$TAG.json 2> gpurun_out/bench_n${N}_b16_$TAG.err
if [ -z "$SKIP_N1" ]; then
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_n1_b32_$TAG.json 2> gpurun_out/bench_n1_b32_$TAG.err
timeout 300 python bench.py --batch 16 --steps 20 --warmup 5 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_n1_b16_$TAG.json 2> gpurun_out/bench_n1_b16_$TAG.err
fi
for f in gpurun_out/bench_n*_$TAG.json; do echo $f; head -c 330 $f; echo; done; tail -5 gpurun_out/bench_n${N}_b32_$TAG.err
