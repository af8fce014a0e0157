"""Multi-GPU plumbing: the batch shards embarrassingly. One process per
GPU (torch.distributed), contiguous slices, no collective on the data path; ONE all-gather of
x (and the status words) only when the caller wants the whole batch on every rank."""
import torch
import torch.distributed as dist


def shard_range(batch, world, rank):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one instance."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(batch, world):
    return [shard_range(batch, world, r)[1] - shard_range(batch, world, r)[0] for r in range(world)]


def all_gather_solutions(x_local, status_local, batch, group=None):
    """x_local [b_r, n], status_local [b_r] (b_r = this rank's shard) -> (x [batch, n],
    status [batch]) on every rank. Works with NCCL (CUDA tensors) and gloo (CPU tensors, used
    by the world_size-2 CPU tests). Ragged shards are padded to the largest shard."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(batch, world)
    bmax, n = max(sizes), x_local.shape[1]
    xpad = torch.zeros((bmax, n), dtype=x_local.dtype, device=x_local.device)
    spad = torch.zeros((bmax,), dtype=status_local.dtype, device=status_local.device)
    xpad[: x_local.shape[0]] = x_local
    spad[: status_local.shape[0]] = status_local
    xs = torch.empty((world * bmax, n), dtype=x_local.dtype, device=x_local.device)
    ss = torch.empty((world * bmax,), dtype=status_local.dtype, device=status_local.device)
    dist.all_gather_into_tensor(xs, xpad, group=group)
    dist.all_gather_into_tensor(ss, spad, group=group)
    if all(s == bmax for s in sizes):
        return xs, ss
    xo = torch.cat([xs[r * bmax: r * bmax + sizes[r]] for r in range(world)])
    so = torch.cat([ss[r * bmax: r * bmax + sizes[r]] for r in range(world)])
    return xo, so
