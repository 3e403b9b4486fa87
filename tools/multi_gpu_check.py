#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun with one rank per GPU:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py
Every rank steps its shard of the env ids (NCCL all-reduce of the statistics at the end, the job's only collective); rank 0
then steps ALL env ids on its own GPU and asserts that the reduced statistics of the sharded job equal those of the
un-sharded one, and that its shard's final boards equal the corresponding slice -- results do not depend on the sharding."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gym_chess_b200 import BatchedChessEnv, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, T = 40000, 700
for opponent, color in (("none", "WHITE"), ("random", "BLACK")):
    off, n = sharding.shard_of(rank, world, N)
    env = BatchedChessEnv(n, opponent=opponent, player_color=color, seed=9, device=local, env_id_offset=off)
    env.step_sampled(T)
    tot = sharding.reduce_stats(env.stats_tensor()).cpu().tolist()
    if rank == 0:
        whole = BatchedChessEnv(N * world, opponent=opponent, player_color=color, seed=9, device=local)
        whole.step_sampled(T)
        ref = whole.stats_tensor().cpu().tolist()
        assert tot == ref, (tot, ref)
        assert torch.equal(env.observe(), whole.observe()[:n])
        print("%s/%s: %d ranks x %d envs x %d steps == one env set of %d (episodes %d, NCCL all-reduce of 16 counters)" % (
            opponent, color, world, n, T, N * world, ref[2]), flush=True)
    dist.barrier()
dist.destroy_process_group()
