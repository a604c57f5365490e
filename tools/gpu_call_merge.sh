#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python -m pytest tests/test_parity_r2_gpu.py tests/test_kernels_gpu.py tests/test_grad_chain_gpu.py -m gpu -q -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -30 > gpurun_out/pytest_merge_$TAG.log
for o in "level1" "l2_32" "l3_64->128s2" "l4_128->256s2"; do
  ONLY=$o KIND=dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape > gpurun_out/merge_${TAG}_a.log; cat gpurun_out/merge_${TAG}_a.log | cut -c1-200
  C3D_DGRAD_MERGE_MAX_O=0 ONLY=$o KIND=dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape | cut -c1-200
  C3D_DGRAD_MERGE_MAX_O=256 ONLY=$o KIND=dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape | cut -c1-200
done
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-iou --skip-torch-baseline > gpurun_out/bench_merge_$TAG.json 2> gpurun_out/bench_merge_$TAG.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log
tail -8 gpurun_out/pytest_merge_$TAG.log; head -c 1500 gpurun_out/bench_merge_$TAG.json; tail -3 gpurun_out/bench_merge_$TAG.err
