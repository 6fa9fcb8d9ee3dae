#!/bin/bash
TAG=${1:-r03v}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host" > $OUT/${TAG}_pytest_host.log 2>&1; echo "host tests exit $?"; tail -3 $OUT/${TAG}_pytest_host.log
timeout 600 python tools/e2e_probe.py 1048576 > $OUT/${TAG}_e2e_2p20.jsonl 2>&1; cat $OUT/${TAG}_e2e_2p20.jsonl | cut -c1-200
timeout 600 python tools/e2e_probe.py 16777216 > $OUT/${TAG}_e2e_2p24.jsonl 2>&1; cat $OUT/${TAG}_e2e_2p24.jsonl | cut -c1-200
