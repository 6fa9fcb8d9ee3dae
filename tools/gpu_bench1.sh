#!/bin/bash
# driver-like: the default bench line and the reference arm, one GPU
TAG=${1:-r03g}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
j=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g frac %.3f ms/step %.4f e2e %.4g launches %s" % (j["value"], j["roofline"]["frac"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"]))
print(j["clocks"])
x=j.get("extras",{})
for k,v in x.items():
    if k=="configs":
        for c in v: print(" ", c["config"], c["mode"], c["envs_per_gpu"], "%.4g"%c["env_steps_per_s"], "%.3f"%c["hbm_frac"])
    elif isinstance(v,dict): print(" ", k, {a:(("%.4g"%b) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(dict,list,str))})
print("cpu_baseline", j.get("cpu_baseline"))
PY
