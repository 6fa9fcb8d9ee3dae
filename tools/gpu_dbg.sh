#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in 1 2 4; do for p in float32 float64; do echo "== V=$v $p"; timeout 300 python tools/debug/many_mismatch.py dense64_continuous 4100 120 $p $v 2>&1 | tail -12; done; done > $OUT/dbg.txt 2>&1
cat $OUT/dbg.txt
