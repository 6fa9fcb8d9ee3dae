#!/bin/bash
# GPU-box visit for tuning: parity tests on the default build, then time every build variant.
TAG=${1:-tune}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 15 $OUT/${TAG}_pytest.log
timeout 900 python tools/tune.py 16777216 200 2>&1 | tee $OUT/${TAG}_tune.txt
