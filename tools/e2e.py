#!/usr/bin/env python
"""e2e timing of the host-buffer step (pinned buffers): tools/e2e.py [envs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_chess_b200 import BatchedChessEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
env = BatchedChessEnv(N, opponent="none", seed=2)
env.step_sampled(300)
words = torch.empty((8, N), dtype=torch.int32).pin_memory(); words.random_(-2**31, 2**31 - 1)
h_r = torch.empty(N, dtype=torch.int32).pin_memory(); h_d = torch.empty(N, dtype=torch.uint8).pin_memory(); h_f = torch.empty(N, dtype=torch.uint8).pin_memory()
wn, rn, dn, fn = words.numpy().view(np.uint32), h_r.numpy(), h_d.numpy(), h_f.numpy()
for i in range(5): env.step_index_host(wn[i % 8], rn, dn, fn)
torch.cuda.synchronize(); t0 = time.time()
K = 200
for i in range(K): env.step_index_host(wn[i % 8], rn, dn, fn)
torch.cuda.synchronize(); dt = time.time() - t0
print("chunks=%s: %.1f us/step, %.3e env steps/s e2e" % (os.environ.get("GCB_HOST_CHUNKS", "default"), dt / K * 1e6, N * K / dt))
