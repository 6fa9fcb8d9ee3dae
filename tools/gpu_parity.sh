#!/bin/bash
# GPU-box visit for parity evidence: all GPU tests, then the measured parity report of both kernel
# families against the oracle.  usage: bash tools/gpu_parity.sh <tag>
TAG=${1:-parity}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 8 $OUT/${TAG}_pytest.log
timeout 1500 python tools/parity_report.py 65536 1000 > $OUT/${TAG}_parity_report.json 2> $OUT/${TAG}_parity_report.err
echo "parity report exit $?"; tail -n 3 $OUT/${TAG}_parity_report.err
