#!/usr/bin/env python
"""Round-2 timing probe (CUDA events, after warm-up): the sampled run kernel at 20 / 64 / 2000 steps, the single-step
kernels (random words / external actions), step + bit mask fused vs as two kernels, the 65,536-env configuration.
    python tools/r2_perf.py [envs] [burn_in]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 600
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


env = BatchedChessEnv(N, opponent="none", seed=2)
if os.environ.get("R2_PERF_DEPHASE", "0") == "1":
    env.dephase()  # the steady-state mix of game phases bench.py times
# default: the batch stays phase-locked (all envs reset together) -- the A/B runs recorded in DESIGN.md 3.2 were made that way;
# their 640-step figure averages over two whole episodes and so agrees with the phase-spread bench, the 20- and 64-step
# figures belong to one window of plies and only compare builds with each other
env.step_sampled(burn)
for k in (20, 64, 640):
    ms = timed(lambda: env.step_sampled(k), reps=5 if k < 600 else 3)
    print("sampled run %4d steps: %.1f us/step  %.3e env steps/s" % (k, ms / k * 1e3, N * k / ms * 1e3))
w = torch.randint(-2**31, 2**31 - 1, (8, N), dtype=torch.int32, device="cuda")
for i in range(3): env.step_index(w[i])
ms = timed(lambda: [env.step_index(w[i % 8]) for i in range(20)])
print("single-step (index words, device): %.1f us/step  %.3e" % (ms / 20 * 1e3, N * 20 / ms * 1e3))
ms = timed(lambda: [env.step_sampled(1) for i in range(20)])
print("single-step (sampled, 1 step per launch): %.1f us/step" % (ms / 20 * 1e3))
bits = torch.empty((N, 66), dtype=torch.int64, device="cuda")
ms = timed(lambda: [(env.step_index(w[i % 8]), env.legal_bitmask(bits)) for i in range(20)])
print("step + bit mask, two kernels: %.1f us/step  %.3e" % (ms / 20 * 1e3, N * 20 / ms * 1e3))
env.set_mask_output(bits)
ms = timed(lambda: [env.step_index(w[i % 8]) for i in range(20)])
print("step + bit mask, fused: %.1f us/step  %.3e" % (ms / 20 * 1e3, N * 20 / ms * 1e3))
env.set_mask_output(None)
ms = timed(lambda: env.legal_bitmask(bits))
print("bit mask kernel alone: %.1f us" % (ms * 1e3))
for opp, col in (("random", "WHITE"), ("random", "BLACK")):
    bot = BatchedChessEnv(N, opponent=opp, player_color=col, seed=2)
    bot.step_sampled(burn // 2)
    ms = timed(lambda: bot.step_sampled(64))
    print("vs random bot, agent %s: %.1f us/step  %.3e env steps/s (2 plies per step)" % (col, ms / 64 * 1e3, N * 64 / ms * 1e3))
    ms = timed(lambda: [bot.step_index(w[i % 8]) for i in range(20)])
    print("   single-step: %.1f us/step" % (ms / 20 * 1e3))
    bot.close()
small = BatchedChessEnv(65536, opponent="none", seed=2)
small.step_sampled(burn)
ms = timed(lambda: small.step_sampled(200))
print("65,536 envs: %.1f us/step  %.3e env steps/s" % (ms / 200 * 1e3, 65536 * 200 / ms * 1e3))
print(env.stats())
