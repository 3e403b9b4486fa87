#!/usr/bin/env python
"""BASELINE.json configs[4]: repetition-heavy endgames, 1M envs, 512-slot ring: step rate in the scan-heavy regime"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
from gym_chess_b200.boards import endgame_boards

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
env = BatchedChessEnv(N, opponent="none", seed=5, initial_boards=endgame_boards(), moves_max=250, history_cap=512)
env.step_sampled(600)
env.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); env.step_sampled(200); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
s = env.stats()
print("endgames: %d envs, %.3f ms/step, %.3e env steps/s, mean window %.1f, extra table probes per ply %.2f, repetitions %d"
      % (N, ms, N / ms * 1e3, s["hist_window"] / s["plies"], s["hist_scanned"] / s["plies"], s["repetitions"]))
