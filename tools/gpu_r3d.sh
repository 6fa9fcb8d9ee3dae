#!/bin/bash
# the whole GPU suite (no -x: every failure is listed), the dense step_many test five more times, then the A/B
TAG=${1:-r03d}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "step_many_equals_stepwise and dense64" 2>&1 | tail -1; done
timeout 900 python tools/many_ab.py 64 > $OUT/${TAG}_ab.txt 2> $OUT/${TAG}_ab.err; echo "ab exit $?"; cut -c1-200 $OUT/${TAG}_ab.txt
