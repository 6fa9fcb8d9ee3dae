#!/bin/bash
TAG=${1:-r03u}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
j=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g frac %.3f moved %s" % (j["value"], j["roofline"]["frac"], j["roofline"]["bytes_per_env_step_moved_by_this_kernel"]))
print(j["cpu_baseline"].get("parity_check")); print(j["cpu_baseline"].get("c1_single_env"))
PY
