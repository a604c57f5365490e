#!/bin/bash
# lab: where does the persistent forward kernel's tile time go?  (C3D_CONV_DBG: 1 no MMA, 2 no TMA, 4 no B, 8 no A, 16 no epilogue)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SH="l4_256->256@40,fpn_out/rpn_256->256@160,l3_128->128@80,l2_64->64@160,l3_root_448"
run() { echo "=== $*"; env "$@" ONLY="$SH" KIND=fwd ITERS=7 python tools/conv_shapes_bench.py 2>&1 | grep -E '"shape"' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('   %-30s %.3f ms %6.0f TF/s'%(r['shape'],r['fwd_ms'],r['fwd_tflops']))"; }
{
for m in 0 1 2 4 8 16 17 18 20 24; do run C3D_CONV_DBG=$m; done
} 2>&1 | tee gpurun_out/conv_lab.txt
