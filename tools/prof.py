#!/usr/bin/env python
"""Short profiling driver: N envs, burn-in so that game phases are mixed, a few sampled steps, one batched movegen.
Used under ncu (see profiles/README.md); prints plain timings when run alone."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gym_chess_b200 import BatchedChessEnv, _lib  # noqa: E402
from gym_chess_b200._lib import Positions, check  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=524288)
ap.add_argument("--burn-in", type=int, default=600)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--opponent", default="none")
ap.add_argument("--no-dephase", action="store_true", help="leave the batch phase-locked (all envs reset together)")
args = ap.parse_args()

env = BatchedChessEnv(args.envs, opponent=args.opponent, seed=2)
if not args.no_dephase:
    env.dephase()  # the steady-state mix of game phases bench.py measures (86 launches: 43 x (7 sampled steps + masked reset))
env.step_sampled(args.burn_in)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
env.step_sampled(args.steps)
e1.record()
torch.cuda.synchronize()
print("step: %.3f ms/launch, %.3e env steps/s" % (e0.elapsed_time(e1) / args.steps, args.envs * args.steps / e0.elapsed_time(e1) * 1e3))
N = args.envs
info = env.info_tensor()
pl = (info[:, 0] < 0).to(torch.uint8).contiguous()
rt = (info[:, 1] + 2 * info[:, 2] + 4 * info[:, 3] + 8 * info[:, 4]).to(torch.uint8).contiguous()
p = env.positions()
pos = Positions(p.bb01, p.bb23, pl.data_ptr(), rt.data_ptr())
out = torch.empty((N, 144), dtype=torch.int16, device="cuda")
cnt = torch.empty(N, dtype=torch.int32, device="cuda")
L = _lib.lib()
for i in range(4):
    if i == 1:
        torch.cuda.synchronize()
        e0.record()
    check(L.gcb_get_possible_moves(N, pos, 0, 0, out.data_ptr(), 144, cnt.data_ptr(), None, None))
e1.record()
torch.cuda.synchronize()
print("movegen: %.3f ms/launch, %.3e positions/s, mean legal %.2f" % (e0.elapsed_time(e1) / 3, N * 3 / e0.elapsed_time(e1) * 1e3, cnt.float().mean().item()))
print(env.stats())
