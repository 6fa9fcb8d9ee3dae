#!/bin/bash
# One configuration on its own: rate + episode statistics, then ncu --set full of one steady-state
# step launch.  usage: bash tools/gpu_ncu_config.sh <tag> [env_id] [envs] [precision] [dense64|-]
TAG=${1:-cfg}; ENV=${2:-AquaEnv-v0}; ENVS=${3:-4194304}; PREC=${4:-float32}; TAB=${5:-dense64}
OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/config_rate.py $ENV $ENVS $PREC $TAB"
timeout 300 $CMD 100 > $OUT/${TAG}_rate.json 2> $OUT/${TAG}_rate.err; echo "rate exit $?"; cat $OUT/${TAG}_rate.json
# launches: reset, 4 x random_actions, 30 warm-up steps (each maybe stream + tail) -> skip 40 step launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 40 -c 1 -f -o $OUT/${TAG}_step $CMD 20 > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
