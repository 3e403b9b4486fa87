#!/usr/bin/env python
"""end-to-end rate of PipelinedChessEnv for different shard counts: tools/pipe_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import PipelinedChessEnv
N = 524288
for K in (1, 2, 3, 4, 8):
    n = N // K * K
    pipe = PipelinedChessEnv(n, shards=K, opponent="none", seed=2)
    pipe.burn_in(600)
    H = pipe.shard_envs
    w = torch.empty((8, K, H), dtype=torch.int16).pin_memory(); w.random_(-2**15, 2**15 - 1)
    def loop(steps):
        for k in range(K): pipe.send_words(k, src=w[0, k])
        for i in range(steps):
            for k in range(K):
                pipe.recv(k)
                if i + 1 < steps: pipe.send_words(k, src=w[(i + 1) % 8, k])
    loop(5); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(602); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("shards=%d: %.1f us per step of all envs, %.3e env steps/s" % (K, dt / 602 * 1e6, n * 602 / dt))
    pipe.close()
