"""Host-side data-parallel helpers (backend agnostic: NCCL on the GPUs, gloo in the CPU tests).

The path shards over the batch only: rank r owns the contiguous sample range
[r*B/n, (r+1)*B/n) -- the reference's DistributedSampler / per-rank environments (P/data/loader.py:147-152,
M/r2r/env.py:126-134) -- and the single exchange step is the gradient all-reduce that DDP performs
(P/utils/misc.py:52-58).  DDP averages; here the ranks SUM one flat buffer and the 1/world factor is folded into
the optimizer kernel's pre-scale (engine.FlatParams.adamw_step(grad_scale=1/world))."""
import torch


def world_size(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def shard_range(n, rank, world):
    """Contiguous slice of n samples owned by ``rank`` (sizes differ by at most one; equal when world divides n)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank, world):
    """Slice every tensor of a batch (leading dim = samples) down to this rank's share."""
    out = []
    for t in tensors:
        lo, hi = shard_range(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return tuple(out)


def all_reduce_sum_(flat, group=None):
    """In-place SUM of one flat tensor over the ranks; a no-op at world size 1 (like the reference helpers,
    P/utils/distributed.py:86-101).  -> world size (the caller divides, or folds 1/world into the optimizer)."""
    import torch.distributed as dist
    w = world_size(group)
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return w


def flatten_grads(params):
    """One contiguous fp32 vector of the gradients (missing grads count as zeros), and the split sizes."""
    sizes = [p.numel() for p in params]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    return flat, sizes
