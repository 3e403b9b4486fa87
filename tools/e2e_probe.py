#!/usr/bin/env python
"""Which part of the host-buffer step costs what: gcb_env_step_index_host with all / some / none of the output buffers,
against the single-step kernel on device buffers.  tools/e2e_probe.py [envs]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_chess_b200 import BatchedChessEnv, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
env = BatchedChessEnv(N, opponent="none", seed=2)
env.step_sampled(300)
words = torch.empty((8, N), dtype=torch.int32).pin_memory(); words.random_(-2**31, 2**31 - 1)
h_r = torch.empty(N, dtype=torch.int32).pin_memory(); h_d = torch.empty(N, dtype=torch.uint8).pin_memory(); h_f = torch.empty(N, dtype=torch.uint8).pin_memory()
L = _lib.lib(); fn = L.gcb_env_step_index_host; h = env._h
wp = [C.c_void_p(words[i].data_ptr()) for i in range(8)]; rp, dp, fp = C.c_void_p(h_r.data_ptr()), C.c_void_p(h_d.data_ptr()), C.c_void_p(h_f.data_ptr())
K = 300
for name, args in (("reward+done+flags", (rp, dp, fp)), ("reward only", (rp, None, None)), ("no outputs", (None, None, None)),
                   ("reward+done+flags", (rp, dp, fp))):
    for i in range(5): fn(h, wp[i % 8], *args, None)
    torch.cuda.synchronize(); t0 = time.time()
    for i in range(K): fn(h, wp[i % 8], *args, None)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("host step, %-18s: %.1f us/step, %.3e env steps/s" % (name, dt / K * 1e6, N * K / dt))
dw = words.cuda()
dr = torch.empty(N, dtype=torch.int32, device="cuda"); dd = torch.empty(N, dtype=torch.uint8, device="cuda"); df = torch.empty(N, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(K): env.step_index(dw[i % 8])
e1.record(); torch.cuda.synchronize()
print("device step_index (single-step kernel, events): %.1f us/step" % (e0.elapsed_time(e1) / K * 1e3))
torch.cuda.synchronize(); t0 = time.time()
for i in range(K):
    env.step_index(dw[i % 8]); torch.cuda.synchronize()
dt = time.time() - t0
print("device step_index + sync per step (wall): %.1f us/step" % (dt / K * 1e6))
