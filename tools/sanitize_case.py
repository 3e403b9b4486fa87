#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer: every kernel of the library once, all env modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_chess_b200 import BatchedChessEngine, BatchedChessEnv

rng = np.random.RandomState(0)
for opponent, color in (("none", "WHITE"), ("random", "WHITE"), ("random", "BLACK")):
    env = BatchedChessEnv(1000, opponent=opponent, player_color=color, seed=3)
    env.step_sampled(330)
    env.step_index(torch.from_numpy(rng.randint(0, 2 ** 31, size=1000).astype(np.int32)).cuda())
    env.step(torch.zeros(1000, dtype=torch.int32, device="cuda"))
    env.step_host(np.full(1000, 3112, np.int32))
    env.reset(torch.ones(1000, dtype=torch.uint8))
    b, info, legal = env.export_numpy()
    m = env.legal_mask()
    s = env.stats()
    env.close()
eng = BatchedChessEngine()
out, cnt, chk = eng.get_possible_moves(b, info[:, 0].astype(np.int8), info[:, 1:5].astype(np.uint8))
out2, cnt2, _ = eng.get_possible_moves(b, info[:, 0].astype(np.int8), info[:, 1:5].astype(np.uint8), attack=True)
a = np.where(cnt > 0, out[np.arange(1000), 0], 4100).astype(np.int32)
eng.next_state(b, info[:, 0].astype(np.int8), info[:, 1:5].astype(np.uint8), a)
eng.update_state(b, info[:, 1:5].astype(np.uint8))
torch.cuda.synchronize()
print("sanitize case done", s["steps"], int(cnt.sum()), int(cnt2.sum()))
