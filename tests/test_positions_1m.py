"""The fixed 1,048,576-position movegen set (BASELINE.json configs[1]) is committed by RECIPE: rebuilding it with the
oracle's self-play harvest gives the committed SHA-256.  (The GPU suite rebuilds it with the CUDA env's harvest and
byte-compares every move list: tests/test_gpu_parity.py::test_config2_fixed_positions_full_byte_compare.)"""
import numpy as np

from tests.golden import make_positions_1m as mp


def test_recipe_reproduces_the_committed_set():
    b, p, r = mp.build(mp.harvest_oracle)
    assert b.shape == (1 << 20, 64) and b.dtype == np.int8 and p.shape == (1 << 20,) and r.shape == (1 << 20, 4)
    assert mp.digest(b, p, r) == mp.committed_digest()
    crafted = (1 << 20) - mp.SELFPLAY_ENVS * mp.SELFPLAY_STEPS // mp.SELFPLAY_EVERY
    assert 0.08 < crafted / (1 << 20) < 0.12
    assert abs(float((p > 0).mean()) - 0.5) < 0.01
