#!/bin/bash
# ncu --set full of one warm step launch of a small slab.  usage: bash tools/gpu_ncu_small.sh <tag> [envs]
TAG=${1:-small}; ENVS=${2:-65536}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on --cache-control none -k regex:step_ -s 200 -c 1 -f -o $OUT/${TAG}_step python tools/small_slab_probe.py $ENVS > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
