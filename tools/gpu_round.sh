#!/bin/bash
# Full GPU-box visit: all GPU tests, smoke, bench (+reference arm), DQN bench, ncu launch list and
# one full capture of the step kernel.  usage: bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 25 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -n 3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -n 5 $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 600 python tools/dqn_bench.py > $OUT/${TAG}_dqn.json 2> $OUT/${TAG}_dqn.err; echo "dqn exit $?"; cat $OUT/${TAG}_dqn.json; tail -n 3 $OUT/${TAG}_dqn.err
NCU_CMD="python bench.py --steps 6 --warmup 3 --no-extras --e2e-steps 1 --envs-per-gpu 16777216"
timeout 300 $NCU_CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $NCU_CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 300 $NCU_CMD > $OUT/${TAG}_ncu_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 4 -c 2 -f -o $OUT/${TAG}_step $NCU_CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
