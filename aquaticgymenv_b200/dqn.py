"""Batched Double-DQN trainer in torch driving an :class:`AquaVecEnv` on the device.

Re-statement of the reference's agent (main/impl/dqn.py; TensorFlow is not part of this stack) for
BASELINE config 5: same hyper-parameters (dqn.py:17-30), network 5 -> 64 -> 64 -> 3 with ReLU and
Adam 1e-3 (dqn.py:301-314), epsilon-greedy with probability eps/3 per action plus 1-eps on the
argmax (dqn.py:212-228), epsilon decayed once per finished episode (dqn.py:139-140,184), soft target
update after every learner step (dqn.py:294-299), and the same progress files (dqn.py:316-321:
rewards.txt / success.txt via np.savetxt, models/model-%05d).  What changes is the data flow: N
envs step together, actions are sampled on the device, transitions go to a device-resident ring
buffer; the env is consumed through its reset()/step() API unchanged.

Two quirks of the reference are options, default = the textbook form:
  * ``reference_double``: dqn.py:267 takes the argmax over Q(s, .) instead of Q(s', .);
  * ``reference_broadcast_loss``: dqn.py:243-247 broadcasts Q(s,a) [B,1] against the full target
    matrix [B,3], pulling Q(s,a) towards the other actions' current values too.
"""
import os
from pathlib import Path
from time import perf_counter

import numpy as np
import torch

default_hyperparam = {  # dqn.py:17-30
    "BUFFER_SIZE": 50000,
    "EPSILON_FINAL": 0.05,
    "EPSILON_INIT": 1.0,
    "EPSILON_DECAY": 10000,  # either a factor or a number of episodes
    "MINIBATCH_SIZE": 64,
    "GAMMA": 0.98,
    "TAU": 0.005,
    "LOSS_STRATEGY": "double",  # / fixed / standard
    # batched extras (not in the reference)
    "UPDATES_PER_STEP": 1,       # learner steps per vector env step
    "LEARNING_RATE": 0.001,
}


def create_model(n_obs=5, n_actions=3):
    """dqn.py:301-314"""
    return torch.nn.Sequential(torch.nn.Linear(n_obs, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64),
                               torch.nn.ReLU(), torch.nn.Linear(64, n_actions))


class ReplayBuffer(object):
    """Device-resident ring buffer of (s, a, r, s', done); replaces the deque of dqn.py:103."""

    def __init__(self, capacity, n_obs, device):
        self.capacity, self.size, self.head = int(capacity), 0, 0
        self.s = torch.zeros(capacity, n_obs, device=device)
        self.s2 = torch.zeros(capacity, n_obs, device=device)
        self.a = torch.zeros(capacity, dtype=torch.int64, device=device)
        self.r = torch.zeros(capacity, device=device)
        self.d = torch.zeros(capacity, dtype=torch.bool, device=device)

    def push(self, s, a, r, s2, d):
        n = s.shape[0]
        if n >= self.capacity:  # keep the newest `capacity` rows
            s, a, r, s2, d = (t[-self.capacity:] for t in (s, a, r, s2, d))
            n = self.capacity
        idx = (self.head + torch.arange(n, device=s.device)) % self.capacity
        self.s[idx], self.a[idx], self.r[idx], self.s2[idx], self.d[idx] = s, a.long(), r.float(), s2, d
        self.head = (self.head + n) % self.capacity
        self.size = min(self.size + n, self.capacity)

    def sample(self, batch, generator=None):
        idx = torch.randint(0, self.size, (min(batch, self.size),), device=self.s.device, generator=generator)
        return self.s[idx], self.a[idx], self.r[idx], self.s2[idx], self.d[idx]


class BatchedDQN(object):

    def __init__(self, environment, save_directory=None, save_freq=100, hyperparams=None,
                 reference_double=False, reference_broadcast_loss=False, seed=0):
        hp = dict(default_hyperparam)
        hp.update(hyperparams or {})
        if hp["LOSS_STRATEGY"] not in ("double", "fixed", "standard"):
            raise NotImplementedError("no strategy \"%s\" implemented" % hp["LOSS_STRATEGY"])
        self.hp = hp
        self.env = environment
        if self.env.continuous:
            raise ValueError("the DQN agent drives the discrete AquaEnv (dqn.py:100 reads action_space.n)")
        self.device = self.env.device
        self.n_actions = 3
        self.reference_double, self.reference_broadcast_loss = reference_double, reference_broadcast_loss
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        torch.manual_seed(seed)
        self.q_function = create_model(5, self.n_actions).to(self.device)
        self.q_function_target = create_model(5, self.n_actions).to(self.device)
        self.q_function_target.load_state_dict(self.q_function.state_dict())
        self.optimizer = torch.optim.Adam(self.q_function.parameters(), lr=hp["LEARNING_RATE"])
        self.exp_buffer = ReplayBuffer(hp["BUFFER_SIZE"], 5, self.device)
        self.savepath = Path(save_directory) if save_directory else None
        if self.savepath:
            (self.savepath / "models").mkdir(parents=True, exist_ok=True)
        self.save_freq = save_freq
        self.episode = 0
        self.reward_list, self.success_list = [], []
        self.env_steps = self.learner_steps = 0

    # ------------------------------------------------------------------ acting
    @torch.no_grad()
    def _epsilon_greedy_best_action(self, state, epsilon):
        """dqn.py:212-228 for all envs: argmax with probability 1 - eps + eps/3, else uniform."""
        greedy = self.q_function(state).argmax(dim=1)
        n = state.shape[0]
        explore = torch.rand(n, device=self.device, generator=self.gen) < epsilon
        rand_a = torch.randint(0, self.n_actions, (n,), device=self.device, generator=self.gen)
        return torch.where(explore, rand_a, greedy)

    # ------------------------------------------------------------------ learning
    def _targets(self, b_state, b_reward, b_next_state, b_done):
        gamma = self.hp["GAMMA"]
        with torch.no_grad():
            strategy = self.hp["LOSS_STRATEGY"]
            if strategy == "double":      # dqn.py:262-272
                pick_from = b_state if self.reference_double else b_next_state
                best_next = self.q_function(pick_from).argmax(dim=1, keepdim=True)
                future = self.q_function_target(b_next_state).gather(1, best_next).squeeze(1)
            elif strategy == "fixed":     # dqn.py:274-282
                future = self.q_function_target(b_next_state).max(dim=1).values
            else:                         # dqn.py:284-292
                future = self.q_function(b_next_state).max(dim=1).values
            return b_reward + (~b_done).float() * gamma * future

    def _improve_network_gradient(self):
        """dqn.py:238-249"""
        b_state, b_action, b_reward, b_next_state, b_done = self.exp_buffer.sample(self.hp["MINIBATCH_SIZE"], self.gen)
        target = self._targets(b_state, b_reward, b_next_state, b_done)
        q = self.q_function(b_state)
        pred = q.gather(1, b_action.unsqueeze(1))           # [B, 1]
        if self.reference_broadcast_loss:
            full = q.detach().clone()
            full.scatter_(1, b_action.unsqueeze(1), target.unsqueeze(1))
            loss = ((pred - full) ** 2).mean()               # [B,1] against [B,3], as dqn.py:246-247
        else:
            loss = ((pred.squeeze(1) - target) ** 2).mean()
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.optimizer.step()
        self.learner_steps += 1
        return loss

    @torch.no_grad()
    def _improve_target_network(self):
        """dqn.py:294-299"""
        tau = self.hp["TAU"]
        for pt, p in zip(self.q_function_target.parameters(), self.q_function.parameters()):
            pt.mul_(1.0 - tau).add_(p, alpha=tau)

    # ------------------------------------------------------------------ training loop
    def train(self, episodes=None, env_steps=None, verbose=0, log_every=200):
        """dqn.py:119-210, vectorised: runs until `episodes` episodes have finished (or
        `env_steps` vector steps).  Returns (reward_list, success_list) like the reference."""
        hp, env = self.hp, self.env
        decay = hp["EPSILON_DECAY"] if hp["EPSILON_DECAY"] < 1 else \
            (hp["EPSILON_FINAL"] / hp["EPSILON_INIT"]) ** (1 / hp["EPSILON_DECAY"])   # dqn.py:139-140
        n = env.num_envs
        env.reset()
        state = env.observe(normalized=True)
        ep_ret = torch.zeros(n, dtype=torch.float64, device=self.device)
        t0, it = perf_counter(), 0
        next_save = self.save_freq
        while (episodes is None or self.episode < episodes) and (env_steps is None or it < env_steps):
            epsilon = max(hp["EPSILON_INIT"] * decay ** self.episode, hp["EPSILON_FINAL"])  # dqn.py:184
            action = self._epsilon_greedy_best_action(state, epsilon)
            flags = env.step_flags(action)
            reward = env.reward
            done = (flags & 1) != 0
            ep_ret += reward.double()
            # the env auto-resets: for a finished env the planes already hold the next episode's
            # first observation, which is never used as s' because `done` masks the bootstrap
            next_state = env.observe(normalized=True)
            self.exp_buffer.push(state, action, reward, next_state, done)
            for _ in range(hp["UPDATES_PER_STEP"]):
                self._improve_network_gradient()
                self._improve_target_network()
            state = next_state
            self.env_steps += n
            it += 1
            if bool(done.any()):
                fin = done.nonzero().squeeze(1)
                self.reward_list.extend(ep_ret[fin].tolist())
                self.success_list.extend(((flags[fin] & 8) != 0).tolist())       # dqn.py:186
                ep_ret[fin] = 0.0
                self.episode = len(self.reward_list)
                if self.savepath and self.episode >= next_save:
                    self.save_progress()
                    next_save = (self.episode // self.save_freq + 1) * self.save_freq
            if verbose and it % log_every == 0:
                last_r, last_s = self.reward_list[-100:], self.success_list[-100:]
                print("step %7d  episodes %8d  mean_last_100 reward %8.2f  success %3.0f%%  epsilon %4.2f  "
                      "env-steps/s %.3g  learner-steps/s %.3g" % (
                          it, self.episode, np.mean(last_r) if last_r else 0.0,
                          100 * np.mean(last_s) if last_s else 0.0, epsilon,
                          self.env_steps / (perf_counter() - t0), self.learner_steps / (perf_counter() - t0)))
        if self.savepath:
            self.save_progress()
        return self.reward_list, self.success_list

    def save_progress(self):
        """dqn.py:316-321 (np.savetxt files the reference's plotters read; the model as a torch
        state dict instead of a TF SavedModel)."""
        np.savetxt(str(self.savepath / "rewards.txt"), self.reward_list)
        np.savetxt(str(self.savepath / "success.txt"), np.array(self.success_list, dtype=np.float64))
        torch.save(self.q_function.state_dict(),
                   str(self.savepath / "models" / ("model-%05d.pt" % (self.episode // self.save_freq))))
