#!/usr/bin/env python
"""time of the possible_actions decode kernel (k_env_legal_list) and of the mask kernel on 524,288 resident envs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
N = 524288
env = BatchedChessEnv(N, opponent="none", seed=2)
env.step_sampled(400)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): env.legal_actions()
torch.cuda.synchronize(); e0.record()
for _ in range(20): lst, cnt = env.legal_actions()
e1.record(); torch.cuda.synchronize()
print("legal_actions (alloc + decode): %.1f us, mean legal %.2f" % (e0.elapsed_time(e1) / 20 * 1e3, cnt.float().mean().item()))
import ctypes as C
from gym_chess_b200 import _lib
L = _lib.lib()
lst = torch.empty((N, 144), dtype=torch.int16, device="cuda"); cnt = torch.empty(N, dtype=torch.int32, device="cuda")
for _ in range(3): L.gcb_env_legal_actions(env._h, lst.data_ptr(), 144, cnt.data_ptr(), None)
torch.cuda.synchronize(); e0.record()
for _ in range(20): L.gcb_env_legal_actions(env._h, lst.data_ptr(), 144, cnt.data_ptr(), None)
e1.record(); torch.cuda.synchronize()
print("k_env_legal_list alone: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
m = torch.empty((N, 4101), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); e0.record()
for _ in range(5): L.gcb_env_legal_mask(env._h, m.data_ptr(), None)
e1.record(); torch.cuda.synchronize()
print("legal_mask (memset 2.1 GB + scatter): %.1f us" % (e0.elapsed_time(e1) / 5 * 1e3))
bits = torch.empty((N, 65), dtype=torch.int64, device="cuda")
for _ in range(3): L.gcb_env_legal_bitmask(env._h, bits.data_ptr(), 65, None)
torch.cuda.synchronize(); e0.record()
for _ in range(20): L.gcb_env_legal_bitmask(env._h, bits.data_ptr(), 65, None)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
np_mean = (env.observe().reshape(N, 64) != 0).sum(1).float().mean().item() / 2
byt = N * (520 + 40 + 8 * np_mean)
print("legal_bitmask (520 B per env): %.1f us, %.0f GB/s algorithmic (write 520 B + read 40 B state + %.1f slots)" % (us, byt / us / 1e3, np_mean))
