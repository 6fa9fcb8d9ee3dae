#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "step_host" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -n 5 $OUT/${TAG}_pytest.log
timeout 300 python tools/e2e_probe.py 1048576 > $OUT/${TAG}_e2e_1M.jsonl 2>&1; cat $OUT/${TAG}_e2e_1M.jsonl
timeout 300 python tools/e2e_probe.py 16777216 > $OUT/${TAG}_e2e_16M.jsonl 2>&1; cat $OUT/${TAG}_e2e_16M.jsonl
timeout 900 python tools/dqn_train.py > $OUT/${TAG}_dqn_train.jsonl 2> $OUT/${TAG}_dqn_train.err; echo "dqn exit $?"
python -c "
import json
for l in open('$OUT/${TAG}_dqn_train.jsonl'):
    j=json.loads(l); print(j['env_steps_per_s'], j['learner_steps_per_s'], j['greedy_success_rate'], j['best_checkpoint_success'], [c['greedy_success'] for c in j['curve']])"
