"""Multi-GPU: one independent env slab per rank, no data-path collective.

Env instances never interact (the obstacle table is read-only), so the batch shards trivially:
rank ``r`` of ``W`` owns the contiguous global env ids ``[base_r, base_r + n_r)``.  Because the
Philox draw streams are keyed by the GLOBAL env id, the per-env trajectories do not depend on
the number of GPUs.  The only communication is an all-reduce (SUM) of the eight int64 episode
counters at report cadence (NCCL on GPU, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

from . import _native as N


def slab(total_envs, rank, world_size):
    """Contiguous split of ``total_envs`` over ``world_size`` ranks, remainder to the first ranks,
    every slab start a multiple of nothing in particular (env ids are just indices).
    Returns ``(env_id_base, n_local)``."""
    total_envs, rank, world_size = int(total_envs), int(rank), int(world_size)
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside [0, %d)" % (rank, world_size))
    q, r = divmod(total_envs, world_size)
    n_local = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, n_local


def make_sharded(total_envs, rank=None, world_size=None, device=None, **env_kwargs):
    """This rank's slab of a ``total_envs`` job as an :class:`AquaVecEnv` (``env_id_base`` set)."""
    from .vec_env import AquaVecEnv
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    base, n_local = slab(total_envs, rank, world_size)
    return AquaVecEnv(n_local, env_id_base=base, device=device, **env_kwargs)


def all_reduce_stats(stats, group=None, async_op=False):
    """SUM the int64[8] counter vector (``AquaVecEnv.stats_tensor()``) over the ranks, in place.
    With ``async_op`` the work handle is returned so the reduce stays off the step's critical
    path."""
    if stats.dtype != torch.int64 or stats.numel() != N.N_STATS:
        raise ValueError("stats must be an int64 tensor with %d entries" % N.N_STATS)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None if async_op else stats
    work = dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return work if async_op else stats
