#!/bin/bash
# ncu --set full of the step kernel only.  usage: bash tools/gpu_ncu.sh <tag> [envs]
TAG=${1:-prof}; ENVS=${2:-16777216}
OUT=gpurun_out; mkdir -p $OUT
NCU_CMD="python bench.py --steps 6 --warmup 3 --no-extras --e2e-steps 0 --envs-per-gpu $ENVS"
timeout 300 $NCU_CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 4 -c 1 -f -o $OUT/${TAG}_step $NCU_CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/${TAG}_ncu_plain.log | cut -c1-200
