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


class FusedAllGather:
    """Solve fused with the all-gather of the solutions over peer memory (include/opensot_b200.h,
    osb_gather_*): the solver kernel of every rank stores each finished x row and status word
    straight into the gather buffer of every GPU of the box (NVLink peer stores, posted while
    the other stacks are still being solved); a stream-ordered device-side barrier replaces the
    NCCL all-gather that would otherwise follow the kernel. torch.distributed is used once, to
    exchange the 64-byte CUDA IPC handles of the buffers.

        fg = FusedAllGather(engine, global_batch)          # every rank, after init_process_group
        fg.solve(d_blobs_of_my_shard)                      # launch + barrier, on the current stream
        x_all, status_all = fg.x_all, fg.status_all        # [global_batch, n], [global_batch]

    The gather buffer has two halves used by step parity: the tensors solve() returns stay valid while
    the NEXT step runs (a faster peer stores into the other half) and are overwritten by the step after
    that. A barrier that timed out (a peer never arrived) is counted on the device; check() turns that
    into an exception at a host synchronisation point of the caller's choice."""

    def __init__(self, engine, global_batch, group=None):
        self.engine = engine
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.lo, self.hi = shard_range(global_batch, self.world, self.rank)
        # every step that can fail locally (allocation, IPC export / open) is followed by an
        # exchange of success flags, so that all ranks raise together instead of one rank
        # leaving the others inside a collective
        err, handle = "", b""
        try:
            handle = engine.gather_create(global_batch, self.rank, self.world, self.lo)
        except Exception as ex:
            err = "rank %d: %s" % (self.rank, ex)
        got = [None] * self.world
        dist.all_gather_object(got, (err, handle), group=group)
        self._raise_if_any([g[0] for g in got])
        try:
            for r, (_, h) in enumerate(got):
                if r != self.rank:
                    engine.gather_connect(r, h)
            self.x_all, self.status_all = engine.gather_tensors()
        except Exception as ex:
            err = "rank %d: %s" % (self.rank, ex)
        errs = [None] * self.world
        dist.all_gather_object(errs, err, group=group)   # also: every rank has opened every buffer before the first peer store
        self._raise_if_any(errs)

    @staticmethod
    def _raise_if_any(errs):
        bad = [e for e in errs if e]
        if bad:
            raise RuntimeError("fused all-gather unavailable: " + "; ".join(bad))

    def solve(self, d_blobs, stream=None, timeout_ms=2000):
        assert d_blobs.shape[0] == self.hi - self.lo, "this rank solves instances [%d, %d)" % (self.lo, self.hi)
        self.engine.solve_device_gather(d_blobs, stream=stream)
        self.engine.gather_barrier(stream=stream, timeout_ms=timeout_ms)
        self.x_all, self.status_all = self.engine.gather_tensors()      # this step's half
        return self.x_all, self.status_all

    def check(self):
        """synchronises with the device and raises if any barrier so far timed out: the gathered rows of that
        step may be stale or partial"""
        n = self.failures()
        if n:
            raise RuntimeError("fused all-gather: %d device-side barrier(s) timed out on rank %d" % (n, self.rank))

    def barrier(self, stream=None, timeout_ms=2000):
        self.engine.gather_barrier(stream=stream, timeout_ms=timeout_ms)

    def failures(self):
        """number of barriers that timed out on this rank (a peer never arrived)"""
        return self.engine.gather_barrier_failures()
