#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "step_host or single_env or callers" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -n 8 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; python -c "
import json; j=json.loads(open('$OUT/${TAG}_bench.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline']['frac'], j['e2e'])"
DQN_ONLY=1 DQN_UPS=2 timeout 900 python tools/dqn_train.py 32000 512 512 > $OUT/${TAG}_dqn_obs_ups2.jsonl 2> $OUT/${TAG}_dqn.err; echo "dqn exit $?"
python -c "
import json
for l in open('$OUT/${TAG}_dqn_obs_ups2.jsonl'):
    j=json.loads(l); print(j['greedy_success_rate'], j['best_checkpoint_success'], [c['greedy_success'] for c in j['curve']])"
DQN_ONLY=1 DQN_UPS=1 timeout 900 python tools/dqn_train.py 48000 1024 1024 > $OUT/${TAG}_dqn_obs_n1024.jsonl 2>> $OUT/${TAG}_dqn.err; echo "dqn exit $?"
python -c "
import json
for l in open('$OUT/${TAG}_dqn_obs_n1024.jsonl'):
    j=json.loads(l); print(j['greedy_success_rate'], j['best_checkpoint_success'], [c['greedy_success'] for c in j['curve']])"
