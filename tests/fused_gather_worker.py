"""Worker of tests/test_gpu_fused_gather.py (run under torch.distributed.run, one process per rank):
the solve fused with the all-gather over peer memory (opensot_b200.distributed.FusedAllGather)
against the oracle on the whole batch and against solve + NCCL all-gather; prints one JSON line
on rank 0. With fewer GPUs than ranks all ranks share cuda:0 (CUDA IPC works between processes
on one device) and the handles travel over gloo."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import opensot_b200 as ob
    from opensot_b200 import synth
    from opensot_b200.distributed import FusedAllGather, all_gather_solutions, shard_range
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    shared = torch.cuda.device_count() < world
    devi = 0 if shared else local
    torch.cuda.set_device(devi)
    dev = torch.device("cuda", devi)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    shape = os.environ.get("OSB_SHAPE", "C29")
    Bg = int(os.environ.get("OSB_BATCH", "1024"))
    steps = int(os.environ.get("OSB_STEPS", "0"))
    prec = os.environ.get("OSB_PRECISION", "f64")
    desc = synth.shape(shape) if prec == "f64" else synth.shape(shape, ob._abi.F32_EPS_FACTOR)
    lo, hi = shard_range(Bg, world, rank)
    eng = ob.BatchedStackSolver(desc, hi - lo, device=devi, precision=prec)
    blobs = synth.fill(desc, 99, 0, Bg, 0.15)           # every rank generates the whole batch, solves its shard
    mine = torch.from_numpy(eng.pack(blobs[lo:hi])).to(dev)
    fg = FusedAllGather(eng, Bg)
    x_all, st_all = fg.solve(mine)
    torch.cuda.synchronize()
    out = {"world": world, "shared_device": shared, "global_batch": Bg, "shape": shape, "precision": prec,
           "kernel": eng.kernel_name(), "barrier_failures": fg.failures()}
    xg = x_all.cpu().numpy().astype(np.float64); sg = st_all.cpu().numpy()
    if os.environ.get("OSB_ORACLE", "1") == "1":
        from oracle import portlib as P
        xp, stp, _ = P.port_solve(desc, blobs)
        good = stp == 0
        out["status_agree"] = bool(np.array_equal(sg == 0, good))
        out["max_err_vs_oracle"] = float(np.abs(xg - xp)[good].max())
    # the same through the unfused path: solve, then all-gather
    tdt = torch.float64 if prec == "f64" else torch.float32
    dx = torch.empty((hi - lo, desc.n), dtype=tdt, device=dev)
    ds = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    eng.solve_device(mine, dx, ds)
    if shared:
        xa, sa = all_gather_solutions(dx.cpu(), ds.cpu(), Bg)
    else:
        xa, sa = all_gather_solutions(dx, ds, Bg)
    out["equal_to_unfused"] = bool(np.array_equal(xa.cpu().numpy().astype(np.float64), xg) and np.array_equal(sa.cpu().numpy(), sg))
    # two halves by step parity: the gathered batch of a step stays intact while the NEXT step (other inputs) runs
    # and lands in the other half; the step after that reuses the first half
    other = torch.from_numpy(eng.pack(synth.fill(desc, 100, lo, hi - lo, 0.15))).to(dev)
    x2, _ = fg.solve(other)
    torch.cuda.synchronize()
    if not shared:
        dist.barrier()
    out["previous_step_intact"] = bool(np.array_equal(x_all.cpu().numpy().astype(np.float64), xg)) and x2.data_ptr() != x_all.data_ptr()
    x3, _ = fg.solve(mine)
    torch.cuda.synchronize()
    out["halves_alternate"] = bool(x3.data_ptr() == x_all.data_ptr() and np.array_equal(x3.cpu().numpy().astype(np.float64), xg))
    fg.check()
    if steps > 0 and not shared:
        def timed(fn):
            for _ in range(5):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        out["ms_solve_only"] = timed(lambda: eng.solve_device(mine, dx, ds))
        out["ms_solve_then_nccl_allgather"] = timed(lambda: (eng.solve_device(mine, dx, ds), all_gather_solutions(dx, ds, Bg)))
        out["ms_fused"] = timed(lambda: fg.solve(mine))
        out["ms_fused_without_barrier"] = timed(lambda: eng.solve_device_gather(mine))
        out["ms_barrier_alone"] = timed(lambda: fg.barrier())
        out["barrier_failures"] = fg.failures()
    ok = out["barrier_failures"] == 0 and out["equal_to_unfused"] and out["previous_step_intact"] and out["halves_alternate"] \
        and out.get("status_agree", True) \
        and out.get("max_err_vs_oracle", 0.0) <= float(os.environ.get("OSB_TOL", "1e-6"))
    flag = torch.tensor([1 if ok else 0])
    if not shared:
        flag = flag.to(dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok_all_ranks"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if out["ok_all_ranks"] else 1)


if __name__ == "__main__":
    main()
