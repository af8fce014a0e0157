"""CPU suite, part 3: the N>1 path (batch sharding + the one all-gather) on world_size 2 with
the gloo backend. The per-rank solve is stood in for by the CPU restatement (test
infrastructure); what is under test is opensot_b200.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opensot_b200.distributed import all_gather_solutions, shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for batch in (1, 2, 7, 4096, 4097, 10240):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = shard_sizes(batch, world)
            assert sum(sizes) == batch and max(sizes) - min(sizes) <= 1


def _worker(rank, world, port_no, batch, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import portlib as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = P.shape("C12")
    lo, hi = shard_range(batch, world, rank)
    blobs = P.synth(d, 5, lo, hi - lo, 0.3)        # each rank generates only its own slice
    x, st, _ = P.port_solve(d, blobs)
    xg, sg = all_gather_solutions(torch.from_numpy(x), torch.from_numpy(st), batch)
    np.save(os.path.join(out_dir, "x_%d.npy" % rank), xg.numpy())
    np.save(os.path.join(out_dir, "s_%d.npy" % rank), sg.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [64, 37])
def test_two_rank_gather_equals_single_process(tmp_path, port, batch):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port_no, batch, str(tmp_path)), nprocs=2, join=True)
    d = port.shape("C12")
    x_all, st_all, _ = port.port_solve(d, port.synth(d, 5, 0, batch, 0.3))
    for r in range(2):
        xg = np.load(tmp_path / ("x_%d.npy" % r)); sg = np.load(tmp_path / ("s_%d.npy" % r))
        assert xg.shape == x_all.shape
        assert np.array_equal(xg, x_all) and np.array_equal(sg, st_all)


class _FakeEngine:
    """stands in for BatchedStackSolver in the host-side protocol of FusedAllGather: records the
    calls, optionally fails at one step on one rank"""

    def __init__(self, rank, fail_at=None):
        self.rank, self.fail_at, self.connected = rank, fail_at, {}

    def gather_create(self, global_batch, rank, world, first):
        if self.fail_at == "create":
            raise RuntimeError("no memory")
        self.first = first
        return bytes([rank]) * 64

    def gather_connect(self, peer, handle):
        if self.fail_at == "connect":
            raise RuntimeError("IPC refused")
        self.connected[peer] = handle

    def gather_tensors(self):
        return torch.zeros(1), torch.zeros(1)


def _fused_worker(rank, world, port_no, fail_rank, fail_at, out_dir):
    from opensot_b200.distributed import FusedAllGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(rank, fail_at if rank == fail_rank else None)
    try:
        fg = FusedAllGather(eng, 37)
        res = "ok %d %d %s" % (fg.lo, fg.hi, sorted((p, h[0]) for p, h in eng.connected.items()))
    except RuntimeError as ex:
        res = "raised: %s" % ex
    with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
        f.write(res)
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_at", [None, "create", "connect"])
def test_fused_gather_handshake_is_collective(tmp_path, fail_at):
    """handles reach the right peers, shards are the contiguous slices, and a failure on ONE rank
    (allocation, IPC open) makes EVERY rank raise instead of leaving the others in a collective"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    mp.spawn(_fused_worker, args=(2, port_no, 1, fail_at, str(tmp_path)), nprocs=2, join=True)
    got = [open(tmp_path / ("r%d.txt" % r)).read() for r in range(2)]
    if fail_at is None:
        assert got[0] == "ok 0 19 [(1, 1)]" and got[1] == "ok 19 37 [(0, 0)]", got
    else:
        assert all(g.startswith("raised: fused all-gather unavailable: rank 1:") for g in got), got
