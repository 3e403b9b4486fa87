#!/usr/bin/env python
"""movegen / list-decode timing on the FIXED 1,048,576-position set (tests/golden/make_positions_1m.py) and on resident envs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv, _lib
from gym_chess_b200._lib import Positions, check
from tests.golden import make_positions_1m as mp
L = _lib.lib()
dev = torch.device("cuda", 0)
boards, players, rights = mp.build(lambda *a: mp.harvest_gpu(*a))
M = len(boards)
d_b, d_p, d_r = (torch.from_numpy(x).to(dev) for x in (boards, players, rights))
bb01 = torch.empty((M, 2), dtype=torch.int64, device=dev); bb23 = torch.empty((M, 2), dtype=torch.int64, device=dev)
pl, rt = torch.empty(M, dtype=torch.uint8, device=dev), torch.empty(M, dtype=torch.uint8, device=dev)
pos = Positions(bb01.data_ptr(), bb23.data_ptr(), pl.data_ptr(), rt.data_ptr())
check(L.gcb_pack(M, d_b.data_ptr(), d_p.data_ptr(), d_r.data_ptr(), pos, None))
lst = torch.empty((M, 144), dtype=torch.int16, device=dev); cnt = torch.empty(M, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for attack in (0, 1):
    for _ in range(3): check(L.gcb_get_possible_moves(M, pos, attack, 0, lst.data_ptr(), 144, cnt.data_ptr(), None, None))
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): check(L.gcb_get_possible_moves(M, pos, attack, 0, lst.data_ptr(), 144, cnt.data_ptr(), None, None))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("movegen attack=%d: %.1f us per 1M positions, %.3e positions/s, mean moves %.2f" % (attack, ms * 1e3, M / ms * 1e3, cnt.float().mean().item()))
N = 524288
env = BatchedChessEnv(N, opponent="none", seed=2)
env.dephase(); env.step_sampled(300)
lst = torch.empty((N, 144), dtype=torch.int16, device=dev); cnt = torch.empty(N, dtype=torch.int32, device=dev)
for _ in range(3): L.gcb_env_legal_actions(env._h, lst.data_ptr(), 144, cnt.data_ptr(), None)
torch.cuda.synchronize(); e0.record()
for _ in range(20): L.gcb_env_legal_actions(env._h, lst.data_ptr(), 144, cnt.data_ptr(), None)
e1.record(); torch.cuda.synchronize()
print("k_env_legal_list: %.1f us per %d envs, mean legal %.2f" % (e0.elapsed_time(e1) / 20 * 1e3, N, cnt.float().mean().item()))
