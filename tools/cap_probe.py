#!/usr/bin/env python
"""sensitivity of the step kernels to the size of the repetition tables (history_cap): tools/cap_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
N = 524288
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for cap in (512, 256, 64, 16, 512):
    env = BatchedChessEnv(N, opponent="none", seed=2, history_cap=cap)
    env.step_sampled(640)
    torch.cuda.synchronize(); e0.record(); env.step_sampled(1280); e1.record(); torch.cuda.synchronize()
    multi = e0.elapsed_time(e1) / 1280 * 1e3
    w = torch.randint(-2**31, 2**31 - 1, (8, N), dtype=torch.int32, device="cuda")
    for i in range(3): env.step_index(w[i])
    torch.cuda.synchronize(); e0.record()
    for i in range(300): env.step_index(w[i % 8])
    e1.record(); torch.cuda.synchronize()
    print("history_cap %4d (table %5.1f GB): multi-step %.1f us/step, single-step %.1f us/step, overflow %d"
          % (cap, N * 2 * cap * 16 / 1e9, multi, e0.elapsed_time(e1) / 300 * 1e3, env.stats()["hist_overflow"]))
    env.close()
