#!/usr/bin/env python
"""One-off soak: bigger randomized differential runs of the CUDA path against the oracle than the test-suite does."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import parity_helpers as ph
from tests.test_gpu_parity import GpuAdapter, _mg
from gym_chess_b200 import BatchedChessEngine

t0 = time.time()
eng = BatchedChessEngine()
rng = np.random.RandomState(123)
b, p, r = ph.crafted_positions(rng, 150000)
n1 = ph.check_movegen_vs_oracle(_mg(eng), b, p, r, attack=False)
n2 = ph.check_movegen_vs_oracle(_mg(eng), b, p, r, attack=True)
hb, hp, hr = ph.harvest_positions(n_envs=200, steps=330, seed=77)
n3 = ph.check_movegen_vs_oracle(_mg(eng), hb, hp, hr, attack=False)
print("movegen: %d + %d + %d moves compared, %.0f s" % (n1, n2, n3, time.time() - t0), flush=True)
for opponent, color, seed in (("none", "WHITE", 101), ("random", "WHITE", 102), ("random", "BLACK", 103)):
    env = GpuAdapter(600, opponent=opponent, player_color=color, seed=seed, auto_reset=True)
    st = ph.check_sampled_vs_oracle(env, opponent, color, seed, 1200, compare_every=100)
    print(opponent, color, "steps", int(st[0]), "episodes", int(st[2]), "mates", int(st[3]), "reps", int(st[4]), "%.0f s" % (time.time() - t0), flush=True)
    env = GpuAdapter(300, opponent=opponent, player_color=color, seed=seed + 10, auto_reset=True)
    ph.check_state_import_vs_oracle(env, opponent, color, seed + 10, np.random.RandomState(seed))
    env = GpuAdapter(300, opponent=opponent, player_color=color, seed=seed + 20, auto_reset=True)
    ph.check_external_actions_vs_oracle(env, opponent, color, seed + 20, 600, np.random.RandomState(seed + 1), True)
    print("  import + external actions ok %.0f s" % (time.time() - t0), flush=True)
# multi-step launches (shared-memory tile, self-play specialisation, env ranges) against single-step launches, long runs
import torch
from gym_chess_b200 import BatchedChessEnv
for opponent, color, seed, N in (("none", "WHITE", 201, 140000), ("random", "WHITE", 202, 20000), ("random", "BLACK", 203, 20000)):
    a = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=seed)
    b = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=seed)
    T = 1500
    for _ in range(T):
        a.step_sampled(1)
    b.step_sampled(T)
    for x, y in zip(a.export_numpy(), b.export_numpy()):
        assert (x == y).all()
    assert a.stats() == b.stats(), (a.stats(), b.stats())
    assert torch.equal(a.legal_bitmask(), b.legal_bitmask())
    print("multi-step == single-step:", opponent, color, N, "envs x", T, "steps, episodes", a.stats()["episodes"], "%.0f s" % (time.time() - t0), flush=True)
print("soak ok")
