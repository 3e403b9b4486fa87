"""The N>1 path on CPU: two gloo ranks each run their shard (host-emulated device rules), all-reduce the statistics,
and the totals equal one un-sharded run -- sharding is invisible and the only collective is the final reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_per_rank, steps, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_chess_b200 import sharding
    from tests.host_emul import emul

    off, n = sharding.shard_of(rank, world, n_per_rank)
    env = emul.EmulEnv(n, opponent="none", seed=9, env_id_offset=off)
    for _ in range(steps):
        env.step_sampled()
    local = torch.from_numpy(env.stats().astype(np.int64))
    total = sharding.reduce_stats(local)
    t = sharding.max_over_ranks(float(rank + 1))
    boards, info, _ = env.export()
    out[rank] = (total.numpy().copy(), boards.copy(), info.copy(), t)
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_unsharded_run():
    from tests.host_emul import emul

    world, n_per_rank, steps = 2, 24, 320
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29541 + os.getpid() % 200, n_per_rank, steps, out), nprocs=world, join=True)
    whole = emul.EmulEnv(world * n_per_rank, opponent="none", seed=9)
    for _ in range(steps):
        whole.step_sampled()
    wb, wi, _ = whole.export()
    ws = whole.stats().astype(np.int64)
    assert (out[0][0] == ws).all() and (out[1][0] == ws).all()
    assert ws[2] > 0  # episodes ended
    assert (np.concatenate([out[0][1], out[1][1]]) == wb).all() and (np.concatenate([out[0][2], out[1][2]]) == wi).all()
    assert out[0][3] == 2.0 and out[1][3] == 2.0
