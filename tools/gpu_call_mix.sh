#!/bin/bash
TAG=${1:-v1}
bash tools/gpu_call_conv.sh $TAG
bash tools/gpu_call_iou.sh $TAG
