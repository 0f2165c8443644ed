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


class PeerAllReduce:
    """The gradient exchange over NVLink peer memory (csrc/peer.cu, include/melogan_b200.h: mg_peer_*): in-place sum over the
    ranks of one node as three kernel launches on the current stream -- no NCCL call and no host synchronisation, so a whole
    data-parallel training cycle can be captured as ONE CUDA graph, and every rank gets bit-identical sums (rank-order
    addition).  Collective constructor: the CUDA IPC handles of the exchange regions travel once through torch.distributed.
    One node, <= 8 ranks, one process per GPU."""

    def __init__(self, max_floats, group=None, device=None):
        import ctypes
        from . import _native
        self._native, self._ct = _native, ctypes
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.world > 8:
            raise ValueError("PeerAllReduce: at most 8 ranks (one NVSwitch node)")
        self._h = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _native.call("mg_peer_create", self.rank, self.world, int(max_floats), ctypes.byref(self._h),
                         ctypes.cast(handle, ctypes.c_void_p))
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        blob = bytes(torch.cat(allh).cpu().tolist())
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        with torch.cuda.device(self.device):
            _native.call("mg_peer_connect", self._h, ctypes.cast(buf, ctypes.c_void_p))
        dist.barrier(group=group)                    # every rank has opened every region before the first exchange
        self.max_floats = int(max_floats)

    def allreduce_sum_(self, flat):
        """In place; returns the factor that turns the sum into the mean (like allreduce_sum_)."""
        if flat.dtype != torch.float32 or not flat.is_contiguous() or not flat.is_cuda:
            raise ValueError("PeerAllReduce: contiguous float32 CUDA tensors")
        with torch.cuda.device(self.device):
            self._native.call("mg_peer_allreduce_sum", self._h, flat.data_ptr(), flat.numel(),
                              torch.cuda.current_stream(self.device).cuda_stream)
        return 1.0 / self.world

    def close(self):
        if self._h:
            self._native.lib().mg_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
