#!/bin/bash
# ncu --set full of the two dominant kernels on the top shape (FPN output / RPN head conv 256->256 3x3 @160^2, batch 32)
mkdir -p gpurun_out
TAG=${1:-v1}
ONLY="fpn_out/rpn_256->256@160" KIND=fwd ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persistent -c 1 -o gpurun_out/ncu_top_fwd_$TAG -f python tools/conv_shapes_bench.py > gpurun_out/ncu_top_fwd_$TAG.log 2>&1
ONLY="fpn_out/rpn_256->256@160" KIND=wgrad ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc -c 1 -o gpurun_out/ncu_top_wgrad_$TAG -f python tools/conv_shapes_bench.py > gpurun_out/ncu_top_wgrad_$TAG.log 2>&1
ls -la gpurun_out/ncu_top_*_$TAG.ncu-rep
