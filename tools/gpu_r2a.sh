#!/bin/bash
# Round-2 first GPU visit: all GPU tests, smoke, step_many probe, variant A/B.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 30 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -n 3 $OUT/${TAG}_smoke.log
timeout 900 python tools/many_probe.py > $OUT/${TAG}_many.jsonl 2> $OUT/${TAG}_many.err; echo "many exit $?"
tail -n 5 $OUT/${TAG}_many.err
timeout 1200 python tools/tune.py > $OUT/${TAG}_tune.txt 2>&1; echo "tune exit $?"
cat $OUT/${TAG}_tune.txt
