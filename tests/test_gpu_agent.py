"""GPU checks of the rows next to the hot path: fused observation /
normaliser output, batched evaluation harness with the reference's baseline policies, and the
batched DQN trainer consuming the env through reset()/step() (BASELINE config 5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_observe_and_normaliser_match_reference_formula():
    import torch
    from aquaticgymenv_b200 import gym_aqua
    from tests.helpers import reference_main
    reference_main()
    from impl import utils                                        # the reference's own, unmodified
    for precision in ("float32", "float64"):
        env = gym_aqua.make("AquaEnv-v1", num_envs=5000, seed=3, precision=precision)
        env.reset()
        for t in range(5):
            env.step_flags(env.random_actions(t))
        raw = env.observe(normalized=False)
        assert raw.dtype == torch.float32 and raw.shape == (5000, 5) and raw.is_contiguous()
        assert torch.equal(raw, env.obs.float().contiguous())
        norm = env.observe(normalized=True).cpu().numpy()
        single = gym_aqua.make("AquaEnv-v0")
        nz = utils.AquaStateNormalizer(single)                    # main/impl/utils.py:15-27
        want = np.concatenate([nz(o) for o in raw.cpu().numpy().astype(np.float64)[:200]])
        assert np.abs(norm[:200] - want).max() < 1e-6
        assert norm.min() >= 0.0 and norm.max() <= 1.0
        single.close()
        env.close()


def test_optimal_policy_reaches_goal_without_obstacles():
    """test_optimal.py:31-37: the bearing controller on the obstacle-free discrete env."""
    from aquaticgymenv_b200 import gym_aqua, policies
    env = gym_aqua.make("AquaEnv-v0", num_envs=4096, seed=1)
    res = policies.evaluate(env, policies.OptimalAquaPolicy(), episodes=20000)
    assert res["episodes"] >= 20000
    assert res["success_rate"] > 0.9, res
    assert res["reward_mean"] > 20.0, res
    env.close()


def test_random_policy_statistics_match_reference_scale():
    """random policy on the default obstacles ends in collision 93-98 % of
    the time, mean episode length of the order of 90 steps."""
    from aquaticgymenv_b200 import gym_aqua, policies
    env = gym_aqua.make("AquaEnv-v1", num_envs=8192, seed=2)
    res = policies.evaluate(env, policies.RandomAquaPolicy(env), episodes=60000)
    assert 0.90 < res["collided"] / res["episodes"] < 0.99, res
    assert 0.01 < res["success_rate"] < 0.10, res
    assert 40 < res["length_mean"] < 130, res
    env.close()


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_batched_dqn_trains_on_the_gpu_env(tmp_path, cuda_graph):
    """dqn_train_policy.py:19-21 flow on the batched env: a short run improves on the random
    policy's success rate and writes the reference's progress files.  Eagerly, and with one CUDA
    graph per vector step (the env's launches captured together with the torch ops)."""
    import torch
    from aquaticgymenv_b200 import dqn, gym_aqua, policies
    env = gym_aqua.make("AquaEnv-v0", num_envs=2048, seed=5)
    agent = dqn.BatchedDQN(env, save_directory=str(tmp_path), save_freq=5000, cuda_graph=cuda_graph,
                           hyperparams={"EPSILON_DECAY": 20000, "BUFFER_SIZE": 200000,
                                        "MINIBATCH_SIZE": 512, "UPDATES_PER_STEP": 2})
    agent.train(env_steps=1500)
    assert agent.env_steps == 1500 * 2048 and agent.learner_steps == 3000
    assert len(agent.reward_list) == len(agent.success_list) > 10000
    res = policies.evaluate(env, policies.GreedyQPolicy(env, agent.q_function), episodes=8000)
    assert res["success_rate"] > 0.25, res        # random policy: ~0.05; the reference reaches 0.94 after 15k episodes
    assert (tmp_path / "rewards.txt").exists() and (tmp_path / "success.txt").exists()
    # the episodes the env counted are the episodes the agent logged (nothing lost in the device log)
    assert env.stats()["episodes"] >= len(agent.reward_list) > 0
    assert agent.exp_buffer.size == min(1500 * 2048, 200000) == int(agent.exp_buffer.size_dev.item())
    env.close()


def test_dqn_graph_mode_reproduces_the_eager_run():
    """The captured vector step must be the eager one: same seeds (the agent's generator in eager
    mode, the default generator under capture, seeded alike and called alike) give the same
    episode log and the same network, bit for bit.  A host value baked into the graph at capture
    time (epsilon, replay cursor, episode count) would show up here."""
    import torch
    from aquaticgymenv_b200 import dqn, gym_aqua
    out = []
    for cuda_graph in (False, True):
        env = gym_aqua.make("AquaEnv-v0", obstacles=True, num_envs=1024, seed=3)
        agent = dqn.BatchedDQN(env, seed=7, cuda_graph=cuda_graph,
                               hyperparams={"EPSILON_DECAY": 3000, "BUFFER_SIZE": 30000, "MINIBATCH_SIZE": 256})
        agent.train(env_steps=300)   # the replay ring wraps (300 * 1024 > 30000), epsilon decays to its floor
        out.append((list(agent.reward_list), list(agent.success_list),
                    [p.detach().clone() for p in agent.q_function.parameters()],
                    [p.detach().clone() for p in agent.q_function_target.parameters()],
                    agent.exp_buffer.s.clone(), env.stats()))
        env.close()
    (r0, s0, w0, t0, b0, st0), (r1, s1, w1, t1, b1, st1) = out
    assert len(r0) > 1000 and r0 == r1 and s0 == s1
    assert all(torch.equal(a, b) for a, b in zip(w0, w1)) and all(torch.equal(a, b) for a, b in zip(t0, t1))
    assert torch.equal(b0, b1) and st0 == st1


@pytest.mark.parametrize("name,env_id,published,published_reward", [
    ("no_obs", "AquaEnv-v0", 0.938, 38.83), ("with_obs", "AquaEnv-v1", 0.667, 25.52)])
def test_shipped_policies_reproduce_the_published_success_rates(name, env_id, published, published_reward):
    """The reference's only published numbers (example_policies/test_results.pickle, report section
    2.3): its two trained policies reach the goal in 93.8 % (no obstacles) and 66.7 % (default
    obstacles) of 1000 test episodes, mean reward 38.83 / 25.52.  Their weights, read from the
    SavedModels (tests/golden/policy_*.npz, tools/make_policy_fixtures.py), replayed greedily
    (main/testing/test_dqn.py:14-22) on the GPU env must land on the same rates: an end-to-end check
    of dynamics, observation, normaliser and termination logic.  The published rates carry
    sqrt(p (1 - p) / 1000) = 0.8 % / 1.5 % of sampling error; 32,768 episodes here."""
    import os
    from aquaticgymenv_b200 import gym_aqua, policies, savedmodel
    from tests.conftest import GOLDEN
    env = gym_aqua.make(env_id, num_envs=8192, seed=11)
    q = savedmodel.load_q_network(os.path.join(GOLDEN, "policy_%s.npz" % name), device=env.device)
    res = policies.evaluate(env, policies.GreedyQPolicy(env, q), episodes=32768)
    assert res["episodes"] == 32768 and res["truncated"] == 0
    assert abs(res["success_rate"] - published) < 0.04, res
    assert abs(res["reward_mean"] - published_reward) < 2.0, res
    env.close()


def test_reference_callers_run_on_the_dropin_env():
    """The reference's own caller code, imported unmodified from its main/ package, against the
    drop-in: Policy.test (main/testing/__init__.py:17-36) with OptimalAquaPolicy
    (test_optimal.py:8-28) and RandomAquaPolicy (test_random.py:7-38), env made through
    gym.make(...) as test_optimal.py:34 / test_random.py:44 do, and aqua_term_string
    (main/impl/utils.py:4-12) on the returned info."""
    from tests.helpers import reference_main
    gym, _ = reference_main()
    from impl.utils import aqua_term_string
    from testing.test_optimal import OptimalAquaPolicy
    from testing.test_random import RandomAquaPolicy
    env = gym.make("AquaEnv-v0", obstacles=False, seed=20261018)   # seeded: the win count below is a fixed number
    assert type(env).__module__.startswith("aquaticgymenv_b200")
    assert env.reward_goal() == 10 and env.reward_collision() == -10 and env.reward_max_steps() == -10  # aqua.py:86-88
    wins = 0
    for _ in range(16):
        tot_rew, info, tot_steps = OptimalAquaPolicy().test(env, render=True)   # render(): no-op shim
        msg = aqua_term_string(info, tot_rew, tot_steps)
        assert sum(info.values()) == 1 and 1 <= tot_steps <= 1001
        wins += info['Termination.success']
        assert ("REACHED GOAL" in msg) == bool(info['Termination.success'])
        if info['Termination.success']:
            assert tot_rew > 10 - 1e-9          # +10 at the goal plus a positive shaping sum
    assert wins >= 11                           # the bearing controller on an obstacle-free map (16 episodes)
    tot_rew, info, tot_steps = RandomAquaPolicy(strategy=0, env_action_space=env.action_space).test(env, render=False)
    assert sum(info.values()) == 1
    obs = env.reset()
    o2, r, d, i = env.step(2)
    assert abs(env.reward_step() - r) < 1e-9 or d                                 # aqua.py:89-90
    # two envs made alike are independent (the reference is unseeded); an explicit seed replays
    other = gym.make("AquaEnv-v0", obstacles=False)
    assert not np.array_equal(other.reset(), env.reset())
    a, b = gym.make("AquaEnv-v0", seed=5), gym.make("AquaEnv-v0", seed=5)
    assert np.array_equal(a.reset(), b.reset()) and np.array_equal(a.step(0)[0], b.step(0)[0])
    for e in (env, other, a, b):
        e.close()
