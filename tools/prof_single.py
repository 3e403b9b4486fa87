#!/usr/bin/env python
"""profiling driver for the single-step kernel (external random words on device): tools/prof_single.py [steps] [fused]
With GCB_SAMPLED_RANGES=1 the first timed launch is k_env_step launch #99 (1 reset + 86 of dephase() + 9 burn-in + 3 warm-up
before it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
N = 524288
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
env = BatchedChessEnv(N, opponent="none", seed=2)
env.dephase()  # the steady-state mix of game phases bench.py measures
env.step_sampled(576)
if len(sys.argv) > 2:
    bits = torch.empty((N, 66), dtype=torch.int64, device="cuda")
    env.set_mask_output(bits)
w = torch.randint(-2**31, 2**31 - 1, (8, N), dtype=torch.int32, device="cuda")
for i in range(3): env.step_index(w[i])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(steps): env.step_index(w[i % 8])
e1.record(); torch.cuda.synchronize()
print("single-step kernel%s: %.1f us/step" % (" + fused bit mask" if len(sys.argv) > 2 else "", e0.elapsed_time(e1) / steps * 1e3))
