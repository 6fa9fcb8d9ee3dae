#!/bin/bash
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "many or rollout or wide or soak" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -n 5 $OUT/${TAG}_pytest.log
timeout 900 python tools/many_vsweep.py > $OUT/${TAG}_vsweep.txt 2>&1; cat $OUT/${TAG}_vsweep.txt
