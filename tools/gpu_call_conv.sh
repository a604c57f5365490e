#!/bin/bash
# conv-kernel gpurun call: kernel parity tests, per-shape table with and without the swapped kernel, one ncu capture.
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -30 > gpurun_out/pytest_conv_$TAG.log
ONLYS=${ONLYS:-"l2_ l3_"}
for o in $ONLYS; do
  ONLY=$o KIND=fwd,dgrad timeout 300 python tools/conv_shapes_bench.py > gpurun_out/conv_swap_${o}$TAG.log 2>&1
  C3D_CONV_NO_SWAP=1 ONLY=$o KIND=fwd,dgrad timeout 300 python tools/conv_shapes_bench.py > gpurun_out/conv_noswap_${o}$TAG.log 2>&1
done
ONLY="l2_64->64" KIND=fwd ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_swap -c 2 -o gpurun_out/ncu_swap_$TAG -f python tools/conv_shapes_bench.py > gpurun_out/ncu_swap_$TAG.log 2>&1
tail -5 gpurun_out/pytest_conv_$TAG.log
for o in $ONLYS; do echo "== swap $o"; grep -E "^\S.*ms" gpurun_out/conv_swap_${o}$TAG.log | head -20; echo "== noswap $o"; grep -E "^\S.*ms" gpurun_out/conv_noswap_${o}$TAG.log | head -20; done
