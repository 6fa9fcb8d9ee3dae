#!/bin/bash
TAG=${1:-r03j}; OUT=gpurun_out; mkdir -p $OUT
for cfg in c2 headline c3 c4; do for K in 64 128 256; do
  timeout 300 python bench.py --config $cfg --steps-per-launch $K --no-extras --e2e-steps 0 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg K=$K', '%.4g'%j['value'], '%.3f'%j['roofline']['frac'], j['config'].get('envs_per_gpu'), 'ms/step %.4f'%j['ms_per_step'])"
done; done 2>&1 | tee $OUT/${TAG}_ksweep.txt
