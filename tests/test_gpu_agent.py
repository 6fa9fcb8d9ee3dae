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


@pytest.mark.parametrize("in_kernel", [False, True])
@pytest.mark.parametrize("cuda_graph", [False, True])
def test_batched_dqn_trains_on_the_gpu_env(tmp_path, cuda_graph, in_kernel):
    """dqn_train_policy.py:19-21 flow on the batched env: a short run improves on the random
    policy's success rate and writes the reference's progress files.  Eagerly, and with one CUDA
    graph per vector step (the env's launches captured together with the torch ops).  in_kernel: the
    acting side (normaliser, online network, epsilon-greedy, step) is one launch of
    aqua_rollout_policy_dev, the network's weights read in place from the trainer's flat parameter."""
    import torch
    from aquaticgymenv_b200 import dqn, gym_aqua, policies
    env = gym_aqua.make("AquaEnv-v0", num_envs=2048, seed=5)
    agent = dqn.BatchedDQN(env, save_directory=str(tmp_path), save_freq=5000, cuda_graph=cuda_graph,
                           in_kernel_acting=in_kernel,
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


@pytest.mark.parametrize("in_kernel", [False, True])
def test_dqn_graph_mode_reproduces_the_eager_run(in_kernel):
    """The captured vector step must be the eager one: same seeds (the agent's generator in eager
    mode, the default generator under capture, seeded alike and called alike) give the same
    episode log and the same network, bit for bit.  A host value baked into the graph at capture
    time (epsilon, replay cursor, episode count) would show up here."""
    import torch
    from aquaticgymenv_b200 import dqn, gym_aqua
    out = []
    for cuda_graph in (False, True):
        env = gym_aqua.make("AquaEnv-v0", obstacles=True, num_envs=1024, seed=3)
        agent = dqn.BatchedDQN(env, seed=7, cuda_graph=cuda_graph, in_kernel_acting=in_kernel,
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


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_in_kernel_q_network_policy(precision):
    """aqua_rollout_policy: the reference's 5-64-64-3 Q-network (dqn.py:301-314) evaluated inside the
    step kernel by the env's own lane.  (1) Its action is the argmax of the network on the normalised
    observation (float64 numpy evaluation of the same weights; steps where the two best outputs are
    closer than 1e-4 are not judged).  (2) The step it takes is aqua_step with that action, bit for
    bit.  (3) K steps in one launch == K launches of one step.  (4) epsilon-greedy explores at the
    requested rate, and the same first_step_index replays the same exploration."""
    import os
    import torch
    from aquaticgymenv_b200 import gym_aqua, savedmodel
    from tests.conftest import GOLDEN
    n, K = 3000, 40
    mk = lambda: gym_aqua.make("AquaEnv-v1", num_envs=n, seed=23, precision=precision)
    a, b, c = mk(), mk(), mk()
    q = savedmodel.load_q_network(os.path.join(GOLDEN, "policy_with_obs.npz"), device=a.device)
    w = a.pack_q_network(q)
    lin = [m for m in q if isinstance(m, torch.nn.Linear)]
    W = [(m.weight.detach().double().cpu().numpy().T, m.bias.detach().double().cpu().numpy()) for m in lin]
    for e in (a, b, c):
        e.reset()
    judged = 0
    for k in range(K):
        obs = a.observe(normalized=True).double().cpu().numpy()
        h = np.maximum(obs @ W[0][0] + W[0][1], 0.0)
        h = np.maximum(h @ W[1][0] + W[1][1], 0.0)
        qv = h @ W[2][0] + W[2][1]
        top = np.sort(qv, axis=1)
        clear = (top[:, -1] - top[:, -2]) > 1e-4
        acts, rew, flg = a.rollout_policy(w, 1, want_actions=True, want_reward=True, want_flags=True)
        got = acts[0, :n].cpu().numpy()
        assert np.array_equal(got[clear], qv.argmax(1)[clear]), k
        assert got.min() >= 0 and got.max() <= 2
        judged += int(clear.sum())
        b.step_flags(acts[0, :n].contiguous())                      # the same step through aqua_step
        assert torch.equal(b.reward, rew[0, :n]) and torch.equal(b.flags, flg[0, :n]), k
    assert judged > 0.99 * n * K
    assert np.array_equal(a.get_state().cpu().numpy(), b.get_state().cpu().numpy()) and a.stats() == b.stats()
    acts_c, rew_c, flg_c = c.rollout_policy(w, K, want_actions=True, want_reward=True, want_flags=True)   # one launch
    assert np.array_equal(a.get_state().cpu().numpy(), c.get_state().cpu().numpy()) and a.stats() == c.stats()
    # exploration: epsilon = 0.3 replaces the greedy action by a uniform one in 30 % of the steps
    d, e = mk(), mk()
    d.reset()
    e.reset()
    greedy = d.rollout_policy(w, 1, first_step_index=500, want_actions=True)[0][0, :n]
    mixed = e.rollout_policy(w, 1, epsilon=0.3, first_step_index=500, want_actions=True)[0][0, :n]
    changed = float((greedy != mixed).float().mean())
    assert abs(changed - 0.3 * 2 / 3) < 0.04, changed
    f = mk()
    f.reset()
    again = f.rollout_policy(w, 1, epsilon=0.3, first_step_index=500, want_actions=True)[0][0, :n]
    assert torch.equal(mixed, again)
    f.close()
    with pytest.raises(Exception):
        gym_aqua.make("AquaContinuousEnv-v1", num_envs=64, seed=1).rollout_policy(w, 1)
    for x in (a, b, c, d, e):
        x.close()


@pytest.mark.parametrize("name,env_id,published", [("no_obs", "AquaEnv-v0", 0.938), ("with_obs", "AquaEnv-v1", 0.667)])
def test_shipped_policies_replayed_inside_the_kernel(name, env_id, published):
    """The same end-to-end check as above with the network evaluated in the kernel: 8192 first
    episodes (every env's first episode: it ends within 1001 steps) from ONE launch of 1001 steps."""
    import os
    import torch
    from aquaticgymenv_b200 import gym_aqua, savedmodel
    from tests.conftest import GOLDEN
    env = gym_aqua.make(env_id, num_envs=8192, seed=11)
    env.reset()
    q = savedmodel.load_q_network(os.path.join(GOLDEN, "policy_%s.npz" % name), device=env.device)
    _, _, flg = env.rollout_policy(env.pack_q_network(q), 1001, want_flags=True)
    flg = flg[:, :8192]
    done = (flg & 1) != 0
    assert bool(done.any(dim=0).all())                      # every env finished its first episode
    first = done.float().argmax(dim=0)                      # step of the first done
    outcome = flg.gather(0, first[None, :])[0]
    success = float(((outcome & 8) != 0).float().mean())
    assert abs(success - published) < 0.04, success
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
