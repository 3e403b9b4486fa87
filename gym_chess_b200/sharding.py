"""Env sharding across the GPUs of one box: envs are independent units, rank r owns the global
env ids [r*n, (r+1)*n), there is NO per-step collective, and the only exchange is one all-reduce (SUM) of the
16-counter episode-statistics vector at the end of a run (NCCL on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_of(rank, world, envs_per_rank):
    """(env_id_offset, num_envs) of `rank`: the global env id enters the Philox counter, so results do not depend on
    how the envs are split over ranks."""
    assert 0 <= rank < world
    return rank * envs_per_rank, envs_per_rank


def reduce_stats(stats, group=None):
    """all-reduce (SUM) of an int64 [16] statistics vector; returns the job-wide totals on every rank."""
    t = stats.clone() if isinstance(stats, torch.Tensor) else torch.as_tensor(stats, dtype=torch.int64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device=None, group=None):
    """max of a python float over ranks (device timings are reported as the max over ranks)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
