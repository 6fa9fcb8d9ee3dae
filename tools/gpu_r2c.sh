#!/bin/bash
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 30 $OUT/${TAG}_pytest.log
timeout 600 python tools/many_probe.py quick > $OUT/${TAG}_many.jsonl 2> $OUT/${TAG}_many.err; echo "many exit $?"
timeout 900 python bench.py --cpu-seconds 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -n 5 $OUT/${TAG}_bench.err
for CFG in "524288 32 4 128 AquaEnv-v0 dense64" "65536 64 1 64"; do
  NAME=$(echo $CFG | tr ' ' '_')
  timeout 300 python tools/many_one.py $CFG > $OUT/${TAG}_one_${NAME}.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_many -s 3 -c 1 -f -o $OUT/${TAG}_many_${NAME} python tools/many_one.py $CFG > $OUT/${TAG}_ncu_${NAME}.log 2>&1
  echo "ncu $NAME exit $?"
done
