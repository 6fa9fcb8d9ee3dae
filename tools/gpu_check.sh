#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, then the ncu launch list and one full capture of
# the step kernel.  Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 1200 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 30 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -n 5 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -n 5 $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cat $OUT/${TAG}_bench_ref.json
NCU_CMD="python bench.py --steps 6 --warmup 3 --no-extras --e2e-steps 1 --envs-per-gpu 8388608"
$NCU_CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $NCU_CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
$NCU_CMD > $OUT/${TAG}_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_ -s 4 -c 2 -f -o $OUT/${TAG}_step $NCU_CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
