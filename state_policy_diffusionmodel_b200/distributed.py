"""Multi-GPU sampling: trajectories are independent, so the batch is sharded by rows across ranks (one process per
GPU, weights replicated) and the only collective is one all_gather of the finished trajectories.
There is nothing to exchange inside the K-step loop."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) slice of `total` rows owned by `rank`; sizes differ by at most one, earlier ranks larger."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    total = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(total, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}, (lo, hi)


def gather_rows(local, total, group=None):
    """all_gather of row-sharded results (ragged shards are padded to the largest shard)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def sharded_sample(sample_fn, batch, group=None, **kw):
    """Runs `sample_fn(local_batch, **kw) -> (b_local, ...)` on this rank's rows and returns all rows on every rank.
    Per-row keyword tensors (`x_T`, and `noise` on its second axis) are sharded the same way."""
    if not dist.is_initialized():
        return sample_fn(batch, **kw)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    total = next(iter(batch.values())).shape[0]
    local, (lo, hi) = shard_batch(batch, rank, world)
    if kw.get("x_T") is not None:
        kw["x_T"] = kw["x_T"][lo:hi]
    if kw.get("noise") is not None:
        kw["noise"] = kw["noise"][:, lo:hi]
    if hi > lo:
        out = sample_fn(local, **kw)
    else:
        out = None
    shape_src = out if out is not None else None
    if shape_src is None:  # a rank without rows still joins the collective
        ref = kw.get("x_T")
        tail = tuple(ref.shape[1:]) if ref is not None else ()
        out = torch.zeros((0,) + tail, device=next(iter(batch.values())).device)
    return gather_rows(out, total, group)


# ---------------------------------------------------------------------------------------------------
# Data-parallel training: replicas, per-GPU batch, ONE all-reduce over the flat gradient buffer.
# ---------------------------------------------------------------------------------------------------
def allreduce_sum_(flat, group=None, buckets=1):
    """In-place summing all-reduce of a flat gradient buffer, optionally in `buckets` contiguous pieces (each piece is
    an independent collective, so a caller can launch them as the backward pass retires them).  Returns 1/world_size,
    the factor the optimizer folds into the update (clip_grad_norm_ and Adam then see the mean gradient)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    n = flat.numel()
    buckets = max(1, min(int(buckets), n))
    step = (n + buckets - 1) // buckets
    for lo in range(0, n, step):
        dist.all_reduce(flat[lo:lo + step], op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_params_(flat, src=0, group=None):
    """Replicas start from rank `src`'s weights (what DDP does at construction)."""
    if dist.is_available() and dist.is_initialized():
        dist.broadcast(flat, src=src, group=group)
    return flat
