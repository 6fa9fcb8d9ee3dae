"""CPU checks of the host-side agent code (the "next" rows of SURVEY.md section 8f): normaliser,
bearing controller and the batched DQN loop against a tiny fake vector env (no GPU needed)."""
import math
import os

import numpy as np
import torch

from aquaticgymenv_b200 import dqn, policies, spaces
from tests.helpers import reference_main


def _ref_optimal_action(state):
    """main/testing/test_optimal.py:10-28, restated with NumPy scalars."""
    norm = lambda a: (a + 2 * np.pi) % (2 * np.pi)
    boat, goal = state[:2], state[3:]
    boat_angle_wrt_zero = norm(state[2] + np.pi / 2)
    goal_angle_wrt_boat = norm(np.arctan2(*(goal - boat)[::-1]))
    angle_diff = goal_angle_wrt_boat - boat_angle_wrt_zero
    if 8 / 180 * np.pi < abs(angle_diff):
        return 0 if angle_diff > 0 else 1
    return 2


def test_optimal_policy_matches_reference_rule():
    rs = np.random.RandomState(0)
    st = np.stack([rs.uniform(0, 100, 500), rs.uniform(0, 100, 500), rs.uniform(-np.pi, np.pi, 500),
                   rs.uniform(0, 100, 500), rs.uniform(0, 100, 500)], axis=1)
    got = policies.OptimalAquaPolicy().get_action(torch.as_tensor(st)).numpy()
    want = np.array([_ref_optimal_action(s) for s in st])
    assert (got == want).mean() > 0.995  # float64 both sides; ties at the threshold aside
    assert set(got.tolist()) == {0, 1, 2}


class _Env:
    observation_space = spaces.Box(np.array([0, 0, -np.pi, 0, 0]), np.array([100, 100, np.pi, 100, 100]),
                                   dtype=np.float64)


def test_reference_normalizer_accepts_the_dropin_spaces():
    """The reference's own AquaStateNormalizer (main/impl/utils.py:15-33, imported unmodified) on the
    drop-in's observation space: the attributes it reads (.high, .low) behave like gym's."""
    reference_main()
    from impl import utils
    nz = utils.AquaStateNormalizer(_Env())
    s = np.array([50.0, 25.0, -np.pi / 2, 100.0, 0.0])
    out = nz(s)
    assert out.shape == (1, 5)
    assert np.allclose(out[0], [0.5, 0.25, 0.25, 1.0, 0.0])     # utils.py:23-27
    assert np.allclose(nz.inverse(out), s)
    assert s[2] == -np.pi / 2                                    # the input is not modified
    assert "REACHED GOAL" in utils.aqua_term_string({'Termination.time': False, 'Termination.collided': False}, 12.5, 40)
    assert "collided" in utils.aqua_term_string({'Termination.time': False, 'Termination.collided': True}, -3.0, 7)


class FakeVecEnv(object):
    """1-D corridor: action 2 moves right, reaching x >= 1 is a success; 20-step time limit."""
    continuous = False
    autoreset = True

    def __init__(self, n):
        self.num_envs, self.device = n, torch.device("cpu")
        self.x = torch.zeros(n)
        self.t = torch.zeros(n, dtype=torch.int64)
        self.reward = torch.zeros(n)

    def reset(self):
        self.x.zero_()
        self.t.zero_()

    def observe(self, normalized=True):
        o = torch.zeros(self.num_envs, 5)
        o[:, 0] = self.x
        return o

    def step_flags(self, a):
        self.x += torch.where(a == 2, 0.25, -0.05)
        self.t += 1
        succ, tout = self.x >= 1.0, self.t >= 20
        flags = torch.where(succ, 9, torch.where(tout, 5, 0)).to(torch.uint8)
        self.reward = torch.where(succ, 10.0, torch.where(tout, -10.0, 0.1))
        done = flags != 0
        self.x[done] = 0.0
        self.t[done] = 0
        return flags


def test_batched_dqn_loop_learns_the_corridor(tmp_path):
    env = FakeVecEnv(32)
    agent = dqn.BatchedDQN(env, save_directory=str(tmp_path), save_freq=50,
                           hyperparams={"EPSILON_DECAY": 200, "BUFFER_SIZE": 4000, "MINIBATCH_SIZE": 64})
    rewards, success = agent.train(episodes=700)
    assert len(rewards) >= 700 and len(rewards) == len(success)
    assert np.mean(success[-200:]) > 0.8                          # it learns to go right
    assert agent.learner_steps > 0 and agent.env_steps == agent.learner_steps * 32
    r = np.loadtxt(os.path.join(str(tmp_path), "rewards.txt"))    # dqn.py:316-321 artifacts
    s = np.loadtxt(os.path.join(str(tmp_path), "success.txt"))
    assert r.shape == s.shape and set(np.unique(s)) <= {0.0, 1.0}
    assert any(f.startswith("model-") for f in os.listdir(os.path.join(str(tmp_path), "models")))


def test_dqn_targets_and_quirk_options():
    torch.manual_seed(0)
    env = FakeVecEnv(4)
    for strategy in ("double", "fixed", "standard"):
        a = dqn.BatchedDQN(env, hyperparams={"LOSS_STRATEGY": strategy})
        s, s2 = torch.rand(8, 5), torch.rand(8, 5)
        r, d = torch.rand(8), torch.tensor([True, False] * 4)
        t = a._targets(s, r, s2, d)
        assert torch.allclose(t[d], r[d])                         # terminal: no bootstrap
        assert t.shape == (8,)
    a = dqn.BatchedDQN(env, reference_double=True, reference_broadcast_loss=True)
    a.exp_buffer.push(torch.rand(16, 5), torch.randint(0, 3, (16,)), torch.rand(16), torch.rand(16, 5),
                      torch.zeros(16, dtype=torch.bool))
    assert torch.isfinite(a._improve_network_gradient())
    w0 = [p.clone() for p in a.q_function_target.parameters()]
    a._improve_target_network()
    assert any(not torch.equal(x, y) for x, y in zip(w0, a.q_function_target.parameters()))
    try:
        dqn.BatchedDQN(env, hyperparams={"LOSS_STRATEGY": "nope"})
        raise AssertionError("bad strategy accepted")
    except NotImplementedError:
        pass


def test_replay_buffer_ring():
    b = dqn.ReplayBuffer(10, 5, "cpu")
    for k in range(4):
        b.push(torch.full((4, 5), float(k)), torch.zeros(4), torch.full((4,), float(k)), torch.zeros(4, 5),
               torch.zeros(4, dtype=torch.bool))
    assert b.size == 10 and b.head == 6
    assert sorted(set(b.r.tolist())) == [1.0, 2.0, 3.0]           # the oldest rows were overwritten
    b.push(torch.ones(25, 5), torch.zeros(25), torch.arange(25.0), torch.zeros(25, 5), torch.zeros(25, dtype=torch.bool))
    assert b.size == 10 and sorted(b.r.tolist()) == list(map(float, range(15, 25)))


def test_savedmodel_reader_and_policy_fixtures():
    """The tensor-bundle reader (aquaticgymenv_b200/savedmodel.py) against the committed fixtures:
    shapes of the 5 -> 64 -> 64 -> 3 network (dqn.py:301-314), and -- where the reference tree is at
    hand -- the fixtures are exactly what the reader extracts from the shipped SavedModels."""
    from aquaticgymenv_b200 import savedmodel
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, model in (("no_obs", "example_no_obs/models/model-00030"), ("with_obs", "example_with_obs/models/model-00032")):
        w = savedmodel.load_weights(os.path.join(golden, "policy_%s.npz" % name))
        assert [k.shape for k, _ in w] == [(5, 64), (64, 64), (64, 3)] and [b.shape for _, b in w] == [(64,), (64,), (3,)]
        assert all(k.dtype == np.float32 and np.isfinite(k).all() for k, _ in w)
        net = savedmodel.load_q_network(os.path.join(golden, "policy_%s.npz" % name))
        x = torch.rand(7, 5)
        h = torch.relu(x @ torch.as_tensor(w[0][0]) + torch.as_tensor(w[0][1]))
        h = torch.relu(h @ torch.as_tensor(w[1][0]) + torch.as_tensor(w[1][1]))
        want = h @ torch.as_tensor(w[2][0]) + torch.as_tensor(w[2][1])
        assert torch.allclose(net(x), want, atol=1e-5)
        src = os.path.join("/root/reference/example_policies", model)
        if os.path.isdir(src):
            for (k0, b0), (k1, b1) in zip(w, savedmodel.load_weights(src)):
                assert np.array_equal(k0, k1) and np.array_equal(b0, b1)


def test_evaluate_counts_a_fixed_quota_of_episodes_per_env():
    """evaluate() keeps the first ceil(episodes / N) episodes of EVERY env: an env with short
    episodes must not crowd out one with long episodes."""
    class TwoSpeedEnv(FakeVecEnv):
        # even envs finish in 2 steps (collision), odd envs in 10 (success)
        @property
        def obs(self):
            return self.observe()

        def step_flags(self, action):
            self.t += 1
            period = torch.where(torch.arange(self.num_envs) % 2 == 0, 2, 10)
            done = self.t >= period
            odd = torch.arange(self.num_envs) % 2 == 1
            self.flags = torch.where(done, torch.where(odd, 9, 3), 0).to(torch.uint8)
            self.reward = torch.where(done, torch.where(odd, 10.0, -10.0), 0.0)
            self.t[done] = 0
            return self.flags

    class Zero(policies.Policy):
        def get_action(self, state):
            return torch.zeros(state.shape[0], dtype=torch.int64)

    res = policies.evaluate(TwoSpeedEnv(8), Zero(), episodes=40)
    assert res["episodes"] == 40 and res["episodes_per_env"] == 5 and res["truncated"] == 0
    assert res["success"] == 20 and res["collided"] == 20          # not 5x more short episodes than long ones
    assert abs(res["length_mean"] - 6.0) < 1e-9 and abs(res["reward_mean"]) < 1e-9


def test_flat_q_network_is_the_reference_network_in_keras_layout():
    """dqn.FlatQNetwork keeps the 4,739 weights of the 5-64-64-3 network (dqn.py:301-314) in ONE flat
    parameter laid out as Keras stores them -- the vector aqua_rollout_policy reads in place.  Same
    function as the Sequential of create_model with those weights; AquaVecEnv.pack_q_network of that
    Sequential gives back the flat vector; a gradient step moves the flat parameter."""
    from aquaticgymenv_b200 import _native
    from aquaticgymenv_b200.vec_env import AquaVecEnv
    torch.manual_seed(3)
    flat = dqn.FlatQNetwork()
    assert flat.theta.numel() == _native.MLP_WEIGHTS == 5 * 64 + 64 + 64 * 64 + 64 + 64 * 3 + 3
    seq = dqn.create_model(5, 3)
    with torch.no_grad():
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        for m, (w, b) in zip(lin, flat._views()):
            m.weight.copy_(w.t())   # torch keeps [out, in], Keras [in, out]
            m.bias.copy_(b)
    x = torch.rand(257, 5)
    assert torch.allclose(flat(x), seq(x), atol=1e-6)
    assert torch.equal(AquaVecEnv.pack_q_network(seq), flat.theta.detach())
    # header and binding agree on the weight count
    hdr = open(_native.HEADER_PATH).read()
    assert "#define AQUA_MLP_WEIGHTS %d" % _native.MLP_WEIGHTS in hdr
    assert "#define AQUA_GRID_DIM %d" % _native.GRID_DIM in hdr
    before = flat.theta.detach().clone()
    opt = torch.optim.Adam(flat.parameters(), lr=1e-2)
    flat(x).pow(2).mean().backward()
    opt.step()
    assert not torch.equal(before, flat.theta.detach())
    try:
        AquaVecEnv.pack_q_network(torch.nn.Sequential(torch.nn.Linear(5, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3)))
        assert False, "another architecture must be refused"
    except ValueError:
        pass
