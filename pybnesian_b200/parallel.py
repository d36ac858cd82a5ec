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


class guard:
    """Collects an exception raised by this rank's share of a sharded call so that the rank still enters the
    collective that follows; `all_reduce_sum(..., error=g.error)` then raises on EVERY rank (the reference raises
    cleanly from its single process; a rank that raised before the collective would leave the others waiting for
    the NCCL timeout).

        with parallel.guard() as g:
            ... local work that may raise ...
        total = parallel.all_reduce_sum(values, ctx, error=g.error)
    """

    def __init__(self):
        self.error = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc is not None and isinstance(exc, Exception) and active():
            self.error = exc
            return True  # swallowed here, re-raised on every rank after the collective
        return False


_ERR_BYTES = 1024


def _raise_everywhere(error, dist, dev):
    """Called on every rank once the reduced error flag is non-zero: the lowest failing rank broadcasts
    (exception class name, message); every rank raises that exception (its own object on the source rank)."""
    import torch
    w, r = dist.get_world_size(_group), dist.get_rank(_group)
    src = torch.tensor([r if error is not None else w], dtype=torch.int64, device=dev)
    dist.all_reduce(src, op=dist.ReduceOp.MIN, group=_group)
    src = int(src.item())
    buf = torch.zeros(_ERR_BYTES, dtype=torch.uint8, device=dev)
    if r == src:
        raw = (type(error).__name__ + "\n" + str(error)).encode("utf-8", "replace")[:_ERR_BYTES]
        buf[: len(raw)] = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=dist.get_global_rank(_group, src) if _group is not None else src, group=_group)
    if r == src:
        raise error
    name, _, msg = bytes(buf.cpu().tolist()).rstrip(b"\0").decode("utf-8", "replace").partition("\n")
    from ._lib import SingularCovarianceData
    cls = {"SingularCovarianceData": SingularCovarianceData, "ValueError": ValueError, "TypeError": TypeError,
           "IndexError": IndexError, "KeyError": KeyError}.get(name, RuntimeError)
    raise cls(msg + " [raised on rank %d]" % src)


def all_reduce_sum(values, ctx=None, error=None):
    """Element-wise sum over ranks of a float64 numpy vector (returned as a new array).  `error` is the exception this
    rank caught while producing `values` (see `guard`), or None; a flag travels with the vector and, if any rank
    failed, every rank raises after the collective."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    if not active():
        if error is not None:
            raise error
        return values
    import torch
    dist = _dist()
    backend = dist.get_backend(_group)
    t = torch.from_numpy(np.append(values.ravel(), 1.0 if error is not None else 0.0))
    dev = torch.device("cpu")
    if backend == "nccl":
        dev = torch.device("cuda", ctx.device if ctx is not None else local_device())
        t = t.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
    t = t.cpu().numpy()
    if t[-1] != 0.0:
        _raise_everywhere(error, dist, dev)
    return t[:-1].reshape(values.shape)


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
