#!/bin/bash
# parity of the new build, then the same-box A/B of the variants
TAG=${1:-r02v}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
timeout 900 python tools/many_ab.py 64 > $OUT/${TAG}_ab.txt 2> $OUT/${TAG}_ab.err; echo "ab exit $?"; cat $OUT/${TAG}_ab.txt
