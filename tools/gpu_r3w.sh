#!/bin/bash
TAG=${1:-r03w}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_agent.py -m gpu -q -x > $OUT/${TAG}_pytest_agent.log 2>&1; echo "agent exit $?"; tail -12 $OUT/${TAG}_pytest_agent.log
DQN_IN_KERNEL=1 timeout 600 python tools/dqn_train.py > $OUT/${TAG}_dqn_train_in_kernel.jsonl 2> $OUT/${TAG}_dqn_train.err; echo "train exit $?"
python - <<PY
import json
for l in open("$OUT/${TAG}_dqn_train_in_kernel.jsonl"):
    j=json.loads(l); print(j["config"][-90:], "success %.3f (ref %.3f) best %.3f env-steps/s %.3g learner/s %.0f" % (j["greedy_success_rate"], j["reference_published_success"], j["best_checkpoint_success"], j["env_steps_per_s"], j["learner_steps_per_s"]))
PY
