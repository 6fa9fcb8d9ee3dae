#!/bin/bash
TAG=${1:-r04j}; OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/many_one.py 1048576 64 0 0 AquaEnv-v1 -"
timeout 120 $CMD > $OUT/${TAG}_head_plain.log 2>&1; echo "plain exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_many -s 3 -c 1 -f -o $OUT/${TAG}_many_head $CMD > $OUT/${TAG}_head_ncu.log 2>&1; echo "ncu exit $?"
