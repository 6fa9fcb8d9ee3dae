#!/bin/bash
# what the driver runs at round end: the GPU tests, smoke(), the default bench and the reference arm
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
python - <<PY
import json
j=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g frac %.3f e2e %.4g launches %s" % (j["value"], j["roofline"]["frac"], j["e2e"]["value"], j["gpu_launches"]))
for c in j["extras"]["configs"]: print(" ", c["config"], c["mode"], c["envs_per_gpu"], "%.4g"%c["env_steps_per_s"], "%.3f"%c["hbm_frac"])
print(j["cpu_baseline"].get("parity_check")); print(j["cpu_baseline"].get("c1_single_env"))
r=json.loads(open("$OUT/${TAG}_bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
PY
