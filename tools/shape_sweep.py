#!/usr/bin/env python
"""Device-resident timing of one launch per stack shape and precision (BASELINE configs 1-3 and the
test shapes), for the kernel the engine picks and, with OSB_KERNEL=generic in the environment,
for the generic kernel. Prints one JSON line per (shape, precision)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="C29:4096,C38c:8192,C35f:4096,C35h:4096,C12:4096,C64c:1024")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tight", type=float, default=0.10)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--prec", default="f64,f32")
    args = ap.parse_args()
    import torch
    import opensot_b200 as ob
    from opensot_b200 import synth, _abi
    dev = torch.device("cuda", 0)
    for item in args.shapes.split(","):
        name, B = item.split(":"); B = int(B)
        for prec in args.prec.split(","):
            desc = synth.shape(name, 2e2 if prec == "f64" else _abi.F32_EPS_FACTOR)
            eng = ob.BatchedStackSolver(desc, B, device=0, precision=prec)
            blobs = eng.pack(synth.fill(desc, args.seed, 0, B, args.tight))
            tdt = torch.float64 if prec == "f64" else torch.float32
            db = torch.from_numpy(blobs).to(dev)
            dx = torch.empty((B, desc.n), dtype=tdt, device=dev)
            ds = torch.empty(B, dtype=torch.int32, device=dev)
            di = torch.empty(B, dtype=torch.int32, device=dev)
            for _ in range(3):
                eng.solve_device(db, dx, ds, d_iters=di)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                eng.solve_device(db, dx, ds, d_iters=di)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            info = eng.info()
            print(json.dumps({"shape": name, "precision": prec, "batch": B, "ms": round(ms, 4),
                              "stacks_per_s": round(B / ms * 1e3), "unsolved": int((ds != 0).sum()),
                              "mean_iters": round(float(di.float().mean()), 2), "max_iters": int(di.max()), "seed": args.seed,
                              "kernel": os.environ.get("OSB_KERNEL", "auto"), "regs": info["kernel_regs"],
                              "smem": info["smem_per_cta"], "ctas_per_sm": info["ctas_per_sm"]}), flush=True)
            eng.close()


if __name__ == "__main__":
    main()
