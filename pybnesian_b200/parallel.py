"""Multi-GPU plumbing: one process per GPU, launched with torchrun.

The path shards without any data-path collective - the training set is replicated, test rows
(logl / slogl) or (candidate, fold) work items (CV scores, hill climbing) or pair tiles (UCV) are
split over the ranks - so the only communication is a SUM all-reduce of a handful of float64
scalars: NCCL over NVLink when the process group is NCCL (GPU tensors), gloo on CPU tensors in
the CPU tests.  Every element of a reduced vector is produced by exactly one rank (the others
contribute 0.0), so the result does not depend on the reduction order.

The reference has nothing to mirror here: it drives one OpenCL device (opencl/opencl_config.cpp:149-220).
"""
import os

import numpy as np

_group = None
_enabled = None


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def enable(flag=True, group=None):
    """Force sharding on/off (default: on whenever torch.distributed is initialised)."""
    global _enabled, _group
    _enabled, _group = flag, group


def active():
    if _enabled is False:
        return False
    return _dist() is not None and _dist().get_world_size(_group) > 1


def rank():
    return _dist().get_rank(_group) if active() else 0


def world_size():
    return _dist().get_world_size(_group) if active() else 1


def local_device():
    return int(os.environ.get("PBN_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def all_reduce_sum(values, ctx=None):
    """Element-wise sum over ranks of a float64 numpy vector (returned as a new array)."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    if not active():
        return values
    import torch
    dist = _dist()
    backend = dist.get_backend(_group)
    t = torch.from_numpy(values.copy())
    if backend == "nccl":
        dev = torch.device("cuda", ctx.device if ctx is not None else local_device())
        t = t.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
    return t.numpy()


def deal(costs, r=None, w=None):
    """Indices of the work items owned by rank r of w: items are sorted by decreasing cost (ties by index)
    and dealt round-robin, so every rank gets a similar load and every item exactly one owner."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    return [i for pos, i in enumerate(order) if pos % w == r]


def shard_range(n, r=None, w=None):
    """Contiguous slice [begin, end) of n items owned by rank r of w (first n % w ranks get one more)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, extra = divmod(int(n), w)
    begin = r * base + min(r, extra)
    return begin, begin + base + (1 if r < extra else 0)
