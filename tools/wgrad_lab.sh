#!/bin/bash
# lab: where does the wgrad stage time go?  (C3D_WGRAD_DBG: 1 no MMA, 2 no TMA, 4 no B loads, 8 no A loads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SH="l4_256->256@40,fpn_out/rpn_256->256@160,l3_128->128@80,l2_64->64@160,l5_512->512@20"
run() { echo "=== $*"; env "$@" ONLY="$SH" KIND=wgrad ITERS=7 python tools/conv_shapes_bench.py 2>&1 | grep -E '"shape"' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('   %-30s %.3f ms %6.0f TF/s'%(r['shape'],r['wgrad_ms'],r['wgrad_tflops']))"; }
{
run X=0
run C3D_WGRAD_DBG=1
run C3D_WGRAD_DBG=2
run C3D_WGRAD_DBG=4
run C3D_WGRAD_DBG=8
run C3D_WGRAD_DBG=3
run C3D_WGRAD_N128=1
run C3D_WGRAD_N128=1 C3D_WGRAD_DBG=1
run C3D_WGRAD_N128=1 C3D_WGRAD_DBG=2
run C3D_WGRAD_PIX64=1
run C3D_WGRAD_PIX64=1 C3D_WGRAD_DBG=1
run C3D_WGRAD_PIX64=1 C3D_WGRAD_DBG=2
} 2>&1 | tee gpurun_out/wgrad_lab.txt
