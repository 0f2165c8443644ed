"""Data-parallel plumbing of the GAN step: batch sharding and the one exchange step.

Every loss term of the step is a mean over independent samples, so rank r trains on its shard of B/world
rolls and the ranks exchange only the flat gradient buffers: sum-all-reduce, then the 1/world factor is
folded into the fused Adam (grad_scale).  Works on any torch.distributed backend: NCCL over NVLink on the
B200 box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous shard [lo, hi) of n samples for this rank; n must be divisible by world (reference drop_last)."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def shard(t, world, rank, dim=0):
    lo, hi = shard_bounds(t.shape[dim], world, rank)
    return t.narrow(dim, lo, hi - lo)


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks of a flat gradient buffer; returns the factor that turns it into the mean."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def mean_of_rank_means(value, group=None):
    """Loss scalars for logging: mean over ranks of per-rank means (equal shard sizes)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return value
    world = dist.get_world_size(group)
    t = value.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t / world


def broadcast_module_state_(modules, src=0, group=None):
    """Rank-independent parameters and BatchNorm running statistics at start (and for checkpoints)."""
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)
