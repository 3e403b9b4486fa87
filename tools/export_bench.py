#!/usr/bin/env python
"""time of the observation / info export kernel (k_env_export) on 524,288 resident envs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv, _lib
N = 524288
env = BatchedChessEnv(N, opponent="none", seed=2)
env.step_sampled(400)
L = _lib.lib()
b = torch.empty((N, 64), dtype=torch.int8, device="cuda"); inf = torch.empty((N, 16), dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, args in (("boards", (b.data_ptr(), None)), ("info", (None, inf.data_ptr())), ("boards+info", (b.data_ptr(), inf.data_ptr()))):
    for _ in range(3): L.gcb_env_export(env._h, args[0], args[1], None)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): L.gcb_env_export(env._h, args[0], args[1], None)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    byt = N * (40 + (64 if args[0] else 0) + (64 if args[1] else 0))
    print("export %-12s %.1f us, %.0f GB/s algorithmic" % (name, us, byt / us / 1e3))
