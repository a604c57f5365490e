#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-v1}
for o in "fpn_lat" "proj" "rpn_pred"; do
  echo "== default $o"; ONLY=$o KIND=fwd,dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape | cut -c1-230
  echo "== P1X1 $o"; C3D_CONV_P1X1=1 ONLY=$o KIND=fwd,dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape | cut -c1-230
done
echo "== l5 s2 dgrad default / merged"; ONLY="l5_256" KIND=dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape | cut -c1-200
C3D_DGRAD_MERGE_MAX_O=512 ONLY="l5_256" KIND=dgrad timeout 300 python tools/conv_shapes_bench.py 2>&1 | grep shape | cut -c1-200
