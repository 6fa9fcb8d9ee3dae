#!/bin/bash
TAG=${1:-r03q}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_agent.py -m gpu -q -x > $OUT/${TAG}_pytest_agent.log 2>&1; echo "agent exit $?"; tail -15 $OUT/${TAG}_pytest_agent.log
timeout 300 python - <<'PY'
import sys, os, time
sys.path.insert(0, ".")
import torch
from aquaticgymenv_b200 import gym_aqua, savedmodel, policies
for n in (8192, 1 << 17, 1 << 20):
    env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=3); env.reset()
    q = savedmodel.load_q_network("tests/golden/policy_with_obs.npz", device=env.device)
    w = env.pack_q_network(q)
    K = 64
    env.rollout_policy(w, K); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.rollout_policy(w, K); env.rollout_policy(w, K); e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) * 1e-3 / 2
    # the torch path: observe -> network -> argmax -> step, per env-step
    pol = policies.GreedyQPolicy(env, q)
    for _ in range(3): env.step_flags(pol.get_action(None))
    torch.cuda.synchronize(); s.record()
    for _ in range(20): env.step_flags(pol.get_action(None))
    e.record(); torch.cuda.synchronize()
    t2 = s.elapsed_time(e) * 1e-3 / 20
    print("envs %8d  in-kernel policy %.4g env-steps/s   torch observe+network+argmax+step %.4g env-steps/s" % (n, n * K / t, n / t2))
    env.close()
PY
