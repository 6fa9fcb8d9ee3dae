#!/bin/bash
# ncu --set full of one steady-state aqua_step_many launch per BASELINE workload (library's own geometry), after a
# plain run of the same command; then the launch list of the default bench.  usage: bash tools/gpu_ncu_many.sh <tag>
TAG=${1:-r03i}; OUT=gpurun_out; mkdir -p $OUT
run() {  # name envs env_id table
  CMD="python tools/many_one.py $2 64 0 0 $3 $4"
  timeout 300 $CMD > $OUT/${TAG}_$1_plain.log 2>&1; echo "$1 plain exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_many -s 3 -c 1 -f -o $OUT/${TAG}_many_$1 $CMD > $OUT/${TAG}_$1_ncu.log 2>&1; echo "$1 ncu exit $?"
}
run c2 65536 AquaEnv-v1 -
run head 1048576 AquaEnv-v1 -
run c3 131072 AquaContinuousEnv-v1 -
run c4 524288 AquaEnv-v0 dense64
timeout 600 python bench.py --steps 4 --warmup 3 --no-extras > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-extras > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list exit $?"
