#!/usr/bin/env python
"""How much does the mix of game phases inside a warp cost?  (a) all envs in lockstep from the reset: time of each 64-step
launch over one episode cycle; (b) the same envs after longer and longer runs (phases mix as episodes end early)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
N = 524288
env = BatchedChessEnv(N, opponent="none", seed=2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(steps):
    torch.cuda.synchronize(); e0.record(); env.step_sampled(steps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3
env.step_sampled(4); env.reset()
tot = 0.0
for k in range(5):
    n = 64 if k < 4 else 45
    us = timed(n); tot += us * n
    print("lockstep plies %3d-%3d: %.1f us/step" % (k * 64, k * 64 + n - 1, us))
print("lockstep, one episode cycle (301 steps): %.1f us/step on average" % (tot / 301))
done = 301
for target in (602, 1204, 3010, 6020, 12040, 24080):
    env.step_sampled(target - done); done = target
    us = timed(602); done += 602
    print("after %5d steps: %.1f us/step over the next 602 steps (%.2e env steps/s)" % (target, us, N / us * 1e6))
