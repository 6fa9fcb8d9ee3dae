#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -n 6 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-seconds 2 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
j=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("N=%d value %.4g frac %.3f ms/step %.4f launches %d e2e %.4g" % (j["n_gpus"], j["value"], j["roofline"]["frac"], j["ms_per_step"], j["gpu_launches"], j["e2e"]["value"]))
for k,v in j.get("extras",{}).items(): print(k, v if not isinstance(v,list) else [(r["config"], r["mode"], r["envs_per_gpu"], r["precision"][-2:], "%.4g"%r["env_steps_per_s"], "%.3f"%r["hbm_frac"]) for r in v])
PY
