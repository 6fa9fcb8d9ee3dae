#!/usr/bin/env python
"""Where does aqua_step_many differ from K single steps?  python tools/debug/many_mismatch.py name n K precision many_v"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
name, n, K, precision, many_v = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
os.environ["AQUA_MANY_V"] = many_v
from tests import test_gpu_parity as T
cfg = T._cfg(name, seed=41, base=2000)
a = T._make_gpu(cfg, n, precision, autoreset=True)
b = T._make_gpu(cfg, n, precision, autoreset=True)
a.reset(); b.reset()
tape = T._tape(a, K)
flg = torch.zeros(K, tape.shape[1], dtype=torch.uint8, device="cuda")
rew = torch.zeros(K, tape.shape[1], dtype=a.dtype, device="cuda")
states = []
for k in range(K):
    a.step_flags(tape[k, :n].contiguous())
    flg[k, :n] = a.flags; rew[k, :n] = a.reward
    states.append(a.get_state().cpu().numpy().copy())
K1 = K // 3
r1, f1 = b.step_many(tape[:K1].contiguous())
s1 = b.get_state().cpu().numpy().copy()
r2, f2 = b.step_many(tape[K1:].contiguous())
got_f = torch.cat([f1, f2])[:, :n]; got_r = torch.cat([r1, r2])[:, :n]
bad = (got_f != flg[:, :n]).nonzero().cpu().numpy()
print("flag mismatches", len(bad), "first", bad[:10].tolist())
badr = (got_r != rew[:, :n]).nonzero().cpu().numpy()
print("reward mismatches", len(badr), "first", badr[:10].tolist())
print("state after launch 1 equal:", np.array_equal(s1, states[K1 - 1]), " final equal:", np.array_equal(b.get_state().cpu().numpy(), states[-1]))
if len(badr):
    k, i = badr[0]
    print("first bad step", k, "env", i, "flags want/got", int(flg[k, i]), int(got_f[k, i]), "reward", float(rew[k, i]), float(got_r[k, i]))
    done_before = (flg[:k, i] & 1).nonzero().flatten().cpu().numpy()
    print("dones of this env before:", done_before.tolist())
    d = np.abs(s1 - states[K1 - 1]).max(axis=0)
    print("envs differing after launch 1:", np.nonzero(d > 0)[0][:20].tolist())
print(a.stats()); print(b.stats())
