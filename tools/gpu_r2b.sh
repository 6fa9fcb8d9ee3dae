#!/bin/bash
# Round-2 GPU visit: all GPU tests, step_many probe, bench line, ncu captures of step_many.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 40 $OUT/${TAG}_pytest.log
timeout 600 python tools/many_probe.py quick > $OUT/${TAG}_many.jsonl 2> $OUT/${TAG}_many.err; echo "many exit $?"
timeout 900 python bench.py --cpu-seconds 4 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json; tail -n 5 $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
cat $OUT/${TAG}_bench_ref.json | cut -c1-600
for CFG in "65536 64 1 64" "1048576 32 2 128"; do
  NAME=$(echo $CFG | tr ' ' '_')
  timeout 300 python tools/many_one.py $CFG > $OUT/${TAG}_one_${NAME}.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_many -s 3 -c 1 -f -o $OUT/${TAG}_many_${NAME} python tools/many_one.py $CFG > $OUT/${TAG}_ncu_${NAME}.log 2>&1
  echo "ncu $NAME exit $?"
done
