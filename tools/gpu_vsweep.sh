#!/bin/bash
TAG=${1:-r03f}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/many_vsweep.py AquaEnv-v1 - > $OUT/${TAG}_vsweep_v1.txt 2> $OUT/${TAG}_vsweep.err; echo "exit $?"; cat $OUT/${TAG}_vsweep_v1.txt
timeout 600 python tools/many_vsweep.py AquaContinuousEnv-v1 - > $OUT/${TAG}_vsweep_cont.txt 2>> $OUT/${TAG}_vsweep.err; echo "exit $?"; cat $OUT/${TAG}_vsweep_cont.txt
timeout 600 python tools/many_vsweep.py AquaEnv-v0 dense64 > $OUT/${TAG}_vsweep_dense.txt 2>> $OUT/${TAG}_vsweep.err; echo "exit $?"; cat $OUT/${TAG}_vsweep_dense.txt
