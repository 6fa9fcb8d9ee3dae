#!/bin/bash
# bench.py under torchrun on N GPUs, as the driver launches it (plus the reference arm).
N=${1:-2}
TAG=${2:-r02s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err; echo "bench $N exit $?"
tail -n 3 $OUT/${TAG}_bench_${N}gpu.err
python - <<PY
import json
j=json.loads(open("$OUT/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("N=%d value %.4g frac %.3f ms/step %.4f launches %d e2e %s" % (j["n_gpus"], j["value"], j["roofline"]["frac"], j["ms_per_step"], j["gpu_launches"], j["e2e"]["value"]))
print(j["clocks"])
for k,v in j.get("extras",{}).items(): print(k, v if not isinstance(v,list) else [(r["config"], r["mode"], r["envs_per_gpu"], "%.4g"%r["env_steps_per_s"], "%.3f"%r["hbm_frac"]) for r in v])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $OUT/${TAG}_ref_${N}gpu.json 2> $OUT/${TAG}_ref_${N}gpu.err; echo "ref $N exit $?"; cut -c1-300 $OUT/${TAG}_ref_${N}gpu.json
[ "$N" -gt 2 ] && exit 0
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > $OUT/${TAG}_bench_1gpu_driverlike.json 2>/dev/null; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench_1gpu_driverlike.json').read().strip().splitlines()[-1]); print('N=1 steps=20:', j['value'], j['roofline']['frac'], j['gpu_launches'], j['clocks'])"
