#!/usr/bin/env python
"""Upper bound of regrouping envs by material: the same multiset of steady-state positions imported (a) in env order,
(b) sorted by a material key, then 64 sampled steps timed (one launch; the history windows start empty in both)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv
N = 524288
src = BatchedChessEnv(N, opponent="none", seed=2)
src.step_sampled(1500)
b = src.observe().reshape(N, 64); info = src.info_tensor()
players = info[:, 0].to(torch.int8); rights = info[:, 1:5].to(torch.uint8); mc = info[:, 8].to(torch.int32)
occ = (b != 0).sum(1); pawns = (b.abs() == 6).sum(1); own = torch.where(info[:, 0] > 0, (b > 0).sum(1), (b < 0).sum(1))
sl = ((b.abs() == 2) | (b.abs() == 3) | (b.abs() == 4)).sum(1)
keys = {"env order": None, "pieces": occ, "pawns, then pieces": pawns * 64 + occ, "pawns, sliders, pieces": (pawns * 16 + sl) * 64 + occ,
        "pieces, then pawns": occ * 32 + pawns}
dst = BatchedChessEnv(N, opponent="none", seed=3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, k in keys.items():
    for blockwise in ((False,) if k is None else (False, True)):
        if k is None:
            perm = torch.arange(N, device="cuda")
        elif blockwise:   # sort inside blocks of 128 envs only (what a block of the step kernel could do by itself)
            perm = (torch.argsort(k.reshape(-1, 128), dim=1) + torch.arange(0, N, 128, device="cuda")[:, None]).reshape(-1)
        else:
            perm = torch.argsort(k)
        res = []
        for steps in (16, 64):
            dst.set_state(b[perm], players[perm], rights[perm], mc[perm])
            torch.cuda.synchronize(); e0.record(); dst.step_sampled(steps); e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / steps * 1e3)
        print("%-26s %-10s first 16 steps %.1f us/step, first 64 steps %.1f us/step" % (name, "per block" if blockwise else "global", res[0], res[1]))
