#!/bin/bash
# Quick GPU-box visit: parity tests, then the headline bench line without extras.
# usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$2" ]; then K=(-k "$2"); else K=(); fi
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider "${K[@]}" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 12 $OUT/${TAG}_pytest.log
for i in 1 2; do
timeout 300 python bench.py --no-extras --e2e-steps 0 --steps 300 --warmup 30 > $OUT/${TAG}_bench$i.json 2> $OUT/${TAG}_bench$i.err; echo "bench exit $?"
python - <<PY
import json
j=json.loads(open("$OUT/${TAG}_bench$i.json").read().strip().splitlines()[-1])
print("%.4g env-steps/s frac %.3f ms %.4f" % (j["value"], j["roofline"]["frac"], j["ms_per_step"]), j.get("clocks"))
PY
done
