#!/bin/bash
# N-GPU weak-scaling check only (no single-GPU legs): bash tools/gpu_call_mg_short.sh <tag> <ngpus>
SKIP_N1=1 bash tools/gpu_call_mg.sh "$@"
