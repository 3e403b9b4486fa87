#!/usr/bin/env python
"""single-step and multi-step cost per 524,288 envs as a function of the env count (how much of a single-step launch is
launch-level cost: cold caches, ramp, last wave)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for N in (131072, 262144, 524288, 1048576, 2097152, 4194304):
    env = BatchedChessEnv(N, opponent="none", seed=2, history_cap=64)
    env.dephase(); env.step_sampled(300)
    w = torch.randint(-2**31, 2**31 - 1, (4, N), dtype=torch.int32, device="cuda")
    for i in range(3): env.step_index(w[i])
    torch.cuda.synchronize(); e0.record()
    for i in range(20): env.step_index(w[i % 4])
    e1.record(); torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1) / 20
    torch.cuda.synchronize(); e0.record(); env.step_sampled(128); e1.record(); torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / 128
    print("N=%8d: single-step %.1f us per 524,288 envs (%.1f us per launch), multi-step %.1f us per 524,288 envs" % (N, t1 * 1e3 * 524288 / N, t1 * 1e3, t2 * 1e3 * 524288 / N), flush=True)
    env.close(); del env, w
