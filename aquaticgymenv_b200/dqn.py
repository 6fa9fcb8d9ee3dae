"""Batched Double-DQN trainer in torch driving an :class:`AquaVecEnv` on the device.

Re-statement of the reference's agent (main/impl/dqn.py; TensorFlow is not part of this stack) for
BASELINE config 5: same hyper-parameters (dqn.py:17-30), network 5 -> 64 -> 64 -> 3 with ReLU and
Adam 1e-3 (dqn.py:301-314), epsilon-greedy with probability eps/3 per action plus 1-eps on the
argmax (dqn.py:212-228), epsilon decayed once per finished episode (dqn.py:139-140,184), soft target
update after every learner step (dqn.py:294-299), and the same progress files (dqn.py:316-321:
rewards.txt / success.txt via np.savetxt, models/model-%05d).  What changes is the data flow: N
envs step together, actions are sampled on the device, transitions go to a device-resident ring
buffer; the env is consumed through its reset()/step() API unchanged.  The loop never waits for the
device: the episode counter (which drives epsilon), the replay cursor and the log of finished
episodes live on the device and are read back every few dozen vector steps.  With
``cuda_graph=True`` one whole vector step (act, env step, observe, replay push, learner update, soft
target update, logging) is captured once in a CUDA graph and replayed: the ~80 small launches of a
step otherwise cost more in Python dispatch than on the GPU.

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


class FlatQNetwork(torch.nn.Module):
    """The same network (dqn.py:301-314) with all 4,739 weights in ONE flat parameter laid out as Keras
    stores them (W1[5][64], b1, W2[64][64], b2, W3[64][3], b3; kernels [in][out]): exactly the vector
    ``aqua_rollout_policy`` reads, so the online network of the trainer is evaluated inside the step
    kernel without any repacking.  Initialised like ``create_model`` (torch's Linear defaults)."""

    SHAPES = ((5, 64), (64, 64), (64, 3))

    def __init__(self, n_obs=5, n_actions=3):
        super().__init__()
        if (n_obs, n_actions) != (5, 3):
            raise ValueError("the in-kernel policy is the reference's 5-64-64-3 network")
        ref = create_model(n_obs, n_actions)
        parts = []
        for m in ref:
            if isinstance(m, torch.nn.Linear):
                parts += [m.weight.detach().t().contiguous().reshape(-1), m.bias.detach().reshape(-1)]
        self.theta = torch.nn.Parameter(torch.cat(parts).contiguous())

    def _views(self):
        out, at = [], 0
        for i, o in self.SHAPES:
            w = self.theta[at:at + i * o].view(i, o)
            at += i * o
            b = self.theta[at:at + o]
            at += o
            out.append((w, b))
        return out

    def forward(self, x):
        (w1, b1), (w2, b2), (w3, b3) = self._views()
        h = torch.relu(torch.addmm(b1, x, w1))
        h = torch.relu(torch.addmm(b2, h, w2))
        return torch.addmm(b3, h, w3)


class ReplayBuffer(object):
    """Device-resident ring buffer of (s, a, r, s', done); replaces the deque of dqn.py:103."""

    def __init__(self, capacity, n_obs, device):
        self.capacity, self.size, self.head = int(capacity), 0, 0
        self.s = torch.zeros(capacity, n_obs, device=device)
        self.s2 = torch.zeros(capacity, n_obs, device=device)
        self.a = torch.zeros(capacity, dtype=torch.int64, device=device)
        self.r = torch.zeros(capacity, device=device)
        self.d = torch.zeros(capacity, dtype=torch.bool, device=device)
        # the same cursor on the device, for the capturable push_dev / sample_dev
        self.head_dev = torch.zeros((), dtype=torch.int64, device=device)
        self.size_dev = torch.zeros((), dtype=torch.int64, device=device)

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

    def push_dev(self, s, a, r, s2, d):
        """push() with the cursor read from and advanced on the device (no host value involved, so
        the call can be captured in a CUDA graph).  note_pushed() keeps the host cursor in step."""
        n = s.shape[0]
        if n >= self.capacity:
            s, a, r, s2, d = (t[-self.capacity:] for t in (s, a, r, s2, d))
            n = self.capacity
        idx = (self.head_dev + torch.arange(n, device=s.device)) % self.capacity
        self.s[idx], self.a[idx], self.r[idx], self.s2[idx], self.d[idx] = s, a.long(), r.float(), s2, d
        self.head_dev.copy_((self.head_dev + n) % self.capacity)
        self.size_dev.copy_(torch.clamp(self.size_dev + n, max=self.capacity))

    def note_pushed(self, n):
        n = min(n, self.capacity)
        self.head = (self.head + n) % self.capacity
        self.size = min(self.size + n, self.capacity)

    def sample_dev(self, batch, generator=None):
        """`batch` rows drawn uniformly from the filled part, its size read on the device"""
        u = torch.rand(batch, device=self.s.device, generator=generator)
        idx = torch.clamp((u * self.size_dev).long(), max=self.capacity - 1)
        return self.s[idx], self.a[idx], self.r[idx], self.s2[idx], self.d[idx]


class BatchedDQN(object):

    def __init__(self, environment, save_directory=None, save_freq=100, hyperparams=None,
                 reference_double=False, reference_broadcast_loss=False, seed=0, cuda_graph=False,
                 in_kernel_acting=False):
        """``in_kernel_acting``: the acting side of a vector step (normalise, online network, epsilon-greedy,
        env step: dqn.py:161-165,212-228) is ONE launch of ``aqua_rollout_policy_dev`` -- the network is
        evaluated by each env's own lane inside the step kernel -- instead of observe + three torch
        Linear layers + argmax + random draws + ``aqua_step``.  The exploration draws then come from the
        env's ACTION stream (keyed by env id and step index) instead of the torch generator."""
        hp = dict(default_hyperparam)
        hp.update(hyperparams or {})
        if hp["LOSS_STRATEGY"] not in ("double", "fixed", "standard"):
            raise NotImplementedError("no strategy \"%s\" implemented" % hp["LOSS_STRATEGY"])
        self.hp = hp
        self.env = environment
        if self.env.continuous:
            raise ValueError("the DQN agent drives the discrete AquaEnv (dqn.py:100 reads action_space.n)")
        self.device = self.env.device
        self.cuda_graph = bool(cuda_graph)
        if self.cuda_graph and torch.device(self.device).type != "cuda":
            raise ValueError("cuda_graph=True needs an env on a CUDA device")
        self.n_actions = 3
        self.reference_double, self.reference_broadcast_loss = reference_double, reference_broadcast_loss
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        torch.manual_seed(seed)
        self.in_kernel_acting = bool(in_kernel_acting)
        make = FlatQNetwork if self.in_kernel_acting else create_model
        self.q_function = make(5, self.n_actions).to(self.device)
        self.q_function_target = make(5, self.n_actions).to(self.device)
        self.q_function_target.load_state_dict(self.q_function.state_dict())
        self.optimizer = torch.optim.Adam(self.q_function.parameters(), lr=hp["LEARNING_RATE"],
                                          fused=torch.device(self.device).type == "cuda",
                                          capturable=self.cuda_graph)
        self.exp_buffer = ReplayBuffer(hp["BUFFER_SIZE"], 5, self.device)
        self.savepath = Path(save_directory) if save_directory else None
        if self.savepath:
            (self.savepath / "models").mkdir(parents=True, exist_ok=True)
        self.save_freq = save_freq
        self.episode = 0
        self.reward_list, self.success_list = [], []
        self.env_steps = self.learner_steps = 0
        # device-side bookkeeping (flushed into the lists above by _flush_log)
        self._episodes_dev = torch.zeros((), dtype=torch.int64, device=self.device)
        self._log_ret = self._log_succ = None   # allocated by train(): finished episodes since the last flush
        self._log_count = torch.zeros((), dtype=torch.int64, device=self.device)
        self._state = self._ep_ret = self._graph = None

    # ------------------------------------------------------------------ acting
    @torch.no_grad()
    def _epsilon_greedy_best_action(self, state, epsilon):
        """dqn.py:212-228 for all envs: argmax with probability 1 - eps + eps/3, else uniform.
        `epsilon`: a float or a 0-d device tensor."""
        greedy = self.q_function(state).argmax(dim=1)
        n = state.shape[0]
        explore = torch.rand(n, device=self.device, generator=self._rng()) < epsilon
        rand_a = torch.randint(0, self.n_actions, (n,), device=self.device, generator=self._rng())
        return torch.where(explore, rand_a, greedy)

    def _rng(self):
        """the agent's own generator; under CUDA-graph capture the default one (the only generator
        whose state a graph advances on replay)"""
        return None if self.cuda_graph else self.gen

    @torch.no_grad()
    def _epsilon(self, decay):
        """dqn.py:184 from the device-side episode counter: no host round trip.  `decay`: a 0-d
        float64 device tensor (made once per train() call: no host-to-device copy per step, which
        a CUDA-graph capture would not allow)."""
        hp = self.hp
        eps = hp["EPSILON_INIT"] * torch.pow(decay, self._episodes_dev.double())
        return torch.clamp(eps, min=hp["EPSILON_FINAL"]).float()

    @torch.no_grad()
    def _log_finished(self, done, ep_ret, flags):
        """Append (return, success) of the envs that finished this step to the device log: the
        rows go to consecutive slots behind _log_count, everything else to a trash slot."""
        cap = self._log_ret.shape[0] - 1
        pos = self._log_count + torch.cumsum(done, 0, dtype=torch.int64) - 1
        pos = torch.where(done, torch.clamp(pos, max=cap), torch.full_like(pos, cap))
        self._log_ret.scatter_(0, pos, ep_ret)
        self._log_succ.scatter_(0, pos, (flags & 8) != 0)                       # dqn.py:186
        k = done.sum()
        self._log_count += k
        self._episodes_dev += k

    def _flush_log(self):
        """device log -> reward_list / success_list (one synchronisation)"""
        k = int(self._log_count.item())
        if k:
            self.reward_list.extend(self._log_ret[:k].tolist())
            self.success_list.extend(self._log_succ[:k].tolist())
            self._log_count.zero_()
        self.episode = len(self.reward_list)

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
        """dqn.py:238-249 (one learner step from the host-side replay cursor)"""
        loss = self._learn(device_cursor=False)
        self.learner_steps += 1
        return loss

    def _learn(self, device_cursor):
        if device_cursor:
            batch = self.exp_buffer.sample_dev(self.hp["MINIBATCH_SIZE"], self._rng())
        else:
            batch = self.exp_buffer.sample(self.hp["MINIBATCH_SIZE"], self.gen)
        b_state, b_action, b_reward, b_next_state, b_done = batch
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
        return loss

    @torch.no_grad()
    def _improve_target_network(self):
        """dqn.py:294-299, all parameters in two launches"""
        tau = self.hp["TAU"]
        target, online = list(self.q_function_target.parameters()), list(self.q_function.parameters())
        torch._foreach_mul_(target, 1.0 - tau)
        torch._foreach_add_(target, online, alpha=tau)

    # ------------------------------------------------------------------ training loop
    def _vector_step(self):
        """One step of all envs plus the learner updates: device work only, no host value read, so
        it runs eagerly or as the body of a CUDA graph alike."""
        hp, env = self.hp, self.env
        state = self._state
        if self.in_kernel_acting:
            with torch.no_grad():
                self._eps_dev.copy_(self._epsilon(self._decay).reshape(1))                  # dqn.py:184
                env.rollout_policy_dev(self.q_function.theta.detach(), 1, self._eps_dev, self._step_dev,
                                       self._act_out, self._rew_out, self._flg_out)
                self._step_dev += 1
                n = env.num_envs
                action, reward, flags = self._act_out[0, :n].long(), self._rew_out[0, :n], self._flg_out[0, :n]
        else:
            action = self._epsilon_greedy_best_action(state, self._epsilon(self._decay))    # dqn.py:184
            flags = env.step_flags(action)
            reward = env.reward
        done = (flags & 1) != 0
        self._ep_ret += reward.double()
        # the env auto-resets: for a finished env the planes already hold the next episode's
        # first observation, which is never used as s' because `done` masks the bootstrap
        next_state = env.observe(normalized=True)
        self.exp_buffer.push_dev(state, action, reward, next_state, done)
        for _ in range(hp["UPDATES_PER_STEP"]):
            self._learn(device_cursor=True)
            self._improve_target_network()
        with torch.no_grad():
            self._log_finished(done, self._ep_ret, flags)
            self._ep_ret.mul_((~done).double())
            state.copy_(next_state)

    def train(self, episodes=None, env_steps=None, verbose=0, log_every=200):
        """dqn.py:119-210, vectorised: runs until `episodes` episodes have finished (or
        `env_steps` vector steps).  Returns (reward_list, success_list) like the reference.
        Finished episodes reach the lists (and `self.episode`) in batches, every few vector steps,
        so the `episodes` criterion can overshoot by up to 4 vector steps."""
        hp, env = self.hp, self.env
        decay = hp["EPSILON_DECAY"] if hp["EPSILON_DECAY"] < 1 else \
            (hp["EPSILON_FINAL"] / hp["EPSILON_INIT"]) ** (1 / hp["EPSILON_DECAY"])   # dqn.py:139-140
        n = env.num_envs
        env.reset()
        self._state = env.observe(normalized=True)
        self._ep_ret = torch.zeros(n, dtype=torch.float64, device=self.device)
        # finished episodes are logged on the device and read back every `flush_every` vector steps
        # (at most n per step, so the log cannot overflow in between); + 1 trash slot
        flush_every = max(1, min(64, (1 << 22) // n))
        if episodes is not None:  # the stopping criterion looks at the flushed count: keep it tight
            flush_every = min(flush_every, 4)
        self._log_ret = torch.zeros(flush_every * n + 1, dtype=torch.float64, device=self.device)
        self._log_succ = torch.zeros(flush_every * n + 1, dtype=torch.bool, device=self.device)
        self._log_count.zero_()
        self._episodes_dev.fill_(self.episode)
        self._decay = torch.as_tensor(decay, dtype=torch.float64, device=self.device)
        if self.in_kernel_acting:
            stride = (n + 15) // 16 * 16
            self._eps_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
            self._step_dev = torch.full((1,), env._step_index, dtype=torch.int64, device=self.device)
            self._act_out = torch.zeros(1, stride, dtype=torch.int32, device=self.device)
            self._rew_out = torch.zeros(1, stride, dtype=env.dtype, device=self.device)
            self._flg_out = torch.zeros(1, stride, dtype=torch.uint8, device=self.device)
        self._graph = None
        t0, it = perf_counter(), 0
        next_save = (self.episode // self.save_freq + 1) * self.save_freq

        def bookkeeping():
            nonlocal it, next_save
            self.exp_buffer.note_pushed(n)
            self.env_steps += n
            self.learner_steps += hp["UPDATES_PER_STEP"]
            it += 1
            if it % flush_every == 0 or (verbose and it % log_every == 0):
                self._flush_log()
                if self.savepath and self.episode >= next_save:
                    self.save_progress()
                    next_save = (self.episode // self.save_freq + 1) * self.save_freq
            if verbose and it % log_every == 0:
                last_r, last_s = self.reward_list[-100:], self.success_list[-100:]
                print("step %7d  episodes %8d  mean_last_100 reward %8.2f  success %3.0f%%  epsilon %4.2f  "
                      "env-steps/s %.3g  learner-steps/s %.3g" % (
                          it, self.episode, np.mean(last_r) if last_r else 0.0,
                          100 * np.mean(last_s) if last_s else 0.0, float(self._epsilon(self._decay)),
                          self.env_steps / (perf_counter() - t0), self.learner_steps / (perf_counter() - t0)))

        def more():
            return (episodes is None or self.episode < episodes) and (env_steps is None or it < env_steps)

        if self.cuda_graph:
            # three eager steps on a side stream (allocator and optimizer state settle), then the
            # capture; every replay afterwards is one vector step
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    if more():
                        self._vector_step()
                        bookkeeping()
            torch.cuda.current_stream(self.device).wait_stream(side)
            if more():
                steps_before = env._step_index
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._vector_step()
                env._step_index = steps_before  # the captured call was recorded, not run
            while more():
                self._graph.replay()
                env._step_index += 1
                bookkeeping()
        else:
            while more():
                self._vector_step()
                bookkeeping()
        self._flush_log()
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
