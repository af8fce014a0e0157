#!/usr/bin/env python
"""BASELINE configs[4]: batch sweep 256 -> 65536, 29-DoF 3-level stack, fp64 and fp32, on the
GPUs of this process group (run under torchrun for N > 1; each rank solves batch/N instances).
Prints a markdown table (rank 0) and writes JSON lines to the file given by --out."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batches", default="256,1024,4096,16384,65536")
    ap.add_argument("--modes", default="strong,weak", help="strong: the batch is the GLOBAL batch, sliced over the ranks; "
                                                           "weak: every rank solves the whole batch (global = batch x ranks)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import opensot_b200 as ob
    from opensot_b200 import synth
    from opensot_b200.distributed import shard_range
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    desc = synth.shape("C29")
    rows = []
    for mode, prec, b in [(m, p, int(b)) for m in args.modes.split(",") for p in ("f64", "f32") for b in args.batches.split(",")]:
        if mode == "weak" and world == 1:
            continue                    # identical to the strong line on one GPU
        for gb in [b if mode == "strong" else b * world]:
            lo, hi = shard_range(gb, world, rank)
            B = hi - lo
            if B < 1:
                continue
            eng = ob.BatchedStackSolver(desc, B, device=local, precision=prec)
            base = synth.fill(desc, 11, lo % 8192, min(B, 8192), 0.10)
            reps = -(-B // base.shape[0])
            blobs = eng.pack(np.tile(base, (reps, 1))[:B])
            tdt = torch.float64 if prec == "f64" else torch.float32
            db = torch.from_numpy(blobs).to(dev)
            dx = torch.empty((B, desc.n), dtype=tdt, device=dev)
            ds = torch.empty(B, dtype=torch.int32, device=dev)
            for _ in range(3):
                eng.solve_device(db, dx, ds)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                eng.solve_device(db, dx, ds)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            bad = int((ds != 0).sum().item())
            rows.append({"scaling": mode, "precision": prec, "global_batch": gb, "per_gpu_batch": B, "n_gpus": world, "ms_per_batch": ms,
                         "stacks_per_s": gb / (ms * 1e-3), "qp_solves_per_s": 3 * gb / (ms * 1e-3), "unsolved_rank0": bad})
            del eng, db, dx, ds
    if rank == 0:
        print("| scaling | precision | GPUs | global batch | ms / batch | stacks/s | QP solves/s |")
        print("|---|---|---|---|---|---|---|")
        for r in rows:
            print("| %s | %s | %d | %d | %.3f | %.3g | %.3g |" % (r["scaling"], r["precision"], r["n_gpus"], r["global_batch"], r["ms_per_batch"],
                                                        r["stacks_per_s"], r["qp_solves_per_s"]))
        if args.out:
            with open(args.out, "a") as f:
                for r in rows:
                    f.write(json.dumps(r) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
