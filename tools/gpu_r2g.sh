#!/bin/bash
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -n 25 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/${TAG}_smoke.log | cut -c1-300
timeout 600 python bench.py --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline']['frac'], j['e2e'])"
