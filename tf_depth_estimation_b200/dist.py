"""Batch sharding for the view-synthesis loss: one process per GPU, contiguous batch slices.

Every sample's warp / loss is independent: the only cross-sample coupling is the final mean.
A rank therefore runs the unchanged fused kernel on its slice with loss_scale = B_local / B_global, which makes
its gradients the rank's exact share of the global-batch gradient (a SUM all-reduce of the network gradients
then yields the global gradient), and its three loss scalars are combined with the same weights.  No
data-path collective exists; the only exchange is the 12-byte loss all-reduce below (NCCL on GPUs, gloo in the
CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous [lo, hi) of `batch` samples for `rank`; the first batch % world ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world of %d' % (rank, world))
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_snippets(d, rank, world):
    """Slice every batch-first tensor (or list of tensors) of a snippet dict to this rank's samples."""
    B = d['tgt'].shape[0]
    lo, hi = shard_range(B, rank, world)

    def cut(v):
        if isinstance(v, (list, tuple)):
            return [cut(t) for t in v]
        return v[lo:hi].contiguous()
    return {k: cut(v) for k, v in d.items()}


def local_loss_scale(batch_local, batch_global):
    """loss_scale for ViewSynthesisPlan so that sum-reduced rank gradients equal the global-batch gradient."""
    return float(batch_local) / float(batch_global)


def reduce_losses(local_losses, batch_local, batch_global, group=None):
    """Global-batch (pixel, smooth, exp) from per-rank means: sum over ranks of (B_local / B_global) * mean."""
    t = local_losses.detach().clone() * (float(batch_local) / float(batch_global))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def reduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of a list of gradient tensors as one flat bucket."""
    if not (dist.is_available() and dist.is_initialized()) or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].reshape(t.shape))
        off += t.numel()
    return tensors
