#!/bin/bash
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
NCU_CMD="python bench.py --gpus 1 --steps 6 --warmup 3 --no-extras --e2e-steps 0"
timeout 300 $NCU_CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv $NCU_CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 300 python tools/many_one.py 1048576 64 4 128 > $OUT/${TAG}_one.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_many -s 3 -c 2 -f -o $OUT/${TAG}_headline_many python tools/many_one.py 1048576 64 4 128 > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
