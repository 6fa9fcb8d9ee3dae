#!/bin/bash
# same-box A/B of aquaticgymenv_b200/variants/*.so (tools/many_ab.py); usage: bash tools/gpu_ab.sh <tag> [K]
TAG=${1:-ab}; K=${2:-64}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python tools/many_ab.py $K > $OUT/${TAG}_ab.txt 2> $OUT/${TAG}_ab.err; echo "ab exit $?"; cut -c1-260 $OUT/${TAG}_ab.txt
