"""Multi-GPU plumbing: independent env shards, one process per GPU, no collective on the step path.

Envs are split into contiguous ranges of GLOBAL env ids; the Philox key of an env is
the batch's base seed and the GLOBAL env id as a counter word (``rs_config.first_env_id``), so results do not depend on the split.
``torch.distributed`` is used only around the step loop: a barrier, the max-over-ranks of the device time
and, when a caller wants them in one place, a host-side gather of the per-shard outputs.
"""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of global env ids owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def weak_shard(envs_per_rank, rank):
    """Weak scaling: every rank owns ``envs_per_rank`` envs; returns (first_env_id, n_envs)."""
    return rank * envs_per_rank, envs_per_rank


def max_over_ranks(value, group=None):
    """max of a python float over the ranks (device time of the slowest shard)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def gather_on_host(local, n_total, group=None):
    """All-gather per-shard numpy outputs (first axis = local envs) into global env order on every rank."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((width,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return np.concatenate([p.cpu().numpy()[:hi - lo] for p, (lo, hi) in zip(parts, sizes)])
