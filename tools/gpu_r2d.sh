#!/bin/bash
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/many_ksweep.py 65536 1 64 > $OUT/${TAG}_ksweep_65536.jsonl 2>&1; cat $OUT/${TAG}_ksweep_65536.jsonl
timeout 300 python tools/many_ksweep.py 131072 2 128 > $OUT/${TAG}_ksweep_131072.jsonl 2>&1; cat $OUT/${TAG}_ksweep_131072.jsonl
timeout 300 python tools/many_ksweep.py 1048576 4 128 > $OUT/${TAG}_ksweep_1M.jsonl 2>&1; cat $OUT/${TAG}_ksweep_1M.jsonl
timeout 600 python bench.py --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]); print(j['value'], j['e2e'])"
timeout 900 python tools/dqn_train.py 24000 512 512 > $OUT/${TAG}_dqn_train.jsonl 2> $OUT/${TAG}_dqn_train.err; echo "dqn exit $?"; cut -c1-900 $OUT/${TAG}_dqn_train.jsonl; tail -3 $OUT/${TAG}_dqn_train.err
NCU_CMD="python tools/many_ksweep.py 65536 1 64"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_many -c 300 --csv --log-file $OUT/${TAG}_ksweep_launches.csv $NCU_CMD > $OUT/${TAG}_ncu_ksweep.log 2>&1; echo "ncu exit $?"
