#!/bin/bash
TAG=${1:-r03o}; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import bench
class A: seed = 0
print(bench.c1_single_env(A(), steps=10000))
PY
