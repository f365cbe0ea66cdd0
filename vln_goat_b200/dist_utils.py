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


# --------------------------------------------------------------------------------------
# sharded optimizer step (engine.FlatParams.sharded_step): layout arithmetic + collectives that also run on gloo
# --------------------------------------------------------------------------------------
def shard_layout(padded, n_decay, world, rank):
    """The flat buffer of ``padded`` elements (a multiple of 8 * world) is cut into ``world`` equal shards; rank r owns
    [r*S, (r+1)*S).  -> (S, lo, n_decay_local): shard size, first element, and how many of the shard's leading elements
    lie in the weight-decayed region [0, n_decay) of the whole buffer."""
    if padded % (8 * world):
        raise ValueError("padded size %d is not a multiple of 8 * world (%d)" % (padded, 8 * world))
    S = padded // world
    lo = rank * S
    return S, lo, min(max(n_decay - lo, 0), S)


def tail_pieces(n_decay, numel, S, world):
    """The no-decay tail [n_decay, numel) (fp32 vectors every rank reads directly: biases, LayerNorm) split by owning
    shard -> [(owner rank, a, b)] with a < b."""
    out = []
    for r in range(world):
        a, b = max(n_decay, r * S), min(numel, (r + 1) * S)
        if a < b:
            out.append((r, a, b))
    return out


def _backend(group=None):
    import torch.distributed as dist
    return dist.get_backend(group)


def reduce_scatter_sum(out, inp, group=None):
    """out[S] = this rank's shard of the element-wise SUM of inp[world*S] over the ranks.  NCCL: one reduce-scatter;
    other backends (gloo in the CPU tests has none): all-reduce a copy and keep the local slice."""
    import torch.distributed as dist
    if _backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)
        return
    tmp = inp.clone()
    dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
    S = out.numel()
    r = dist.get_rank(group)
    out.copy_(tmp[r * S:(r + 1) * S])


def all_gather_flat(out, shard, group=None):
    """out[world*S] = concatenation of every rank's shard[S] in rank order."""
    import torch.distributed as dist
    if _backend(group) == "nccl":
        dist.all_gather_into_tensor(out, shard, group=group)
        return
    W = dist.get_world_size(group)
    parts = [torch.empty_like(shard) for _ in range(W)]
    dist.all_gather(parts, shard, group=group)
    out.copy_(torch.cat(parts))


def group_src(r, group=None):
    """global rank of group rank r (what dist.broadcast's ``src`` expects)"""
    import torch.distributed as dist
    return dist.get_global_rank(group, r) if group is not None else r
