#!/usr/bin/env python
"""A/B of two builds of the library (OSB_LIB selects the build, see opensot_b200/engine.py): solves the same seeded
batches, prints the device time of a launch (CUDA events, 40 launches after 5) and writes x, status and iteration counts
to an .npz, so that two builds can be compared bit for bit.
    OSB_LIB=/path/to/other/libopensot_b200.so python tools/ab_dump.py out.npz [shapes]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    shapes = sys.argv[2] if len(sys.argv) > 2 else "C29:4096,C29:16384,C38c:8192,C35f:4096,C35h:4096,C64c:1024"
    import torch
    import opensot_b200 as ob
    from opensot_b200 import synth, _abi
    dev = torch.device("cuda", 0)
    res = {}
    for item in shapes.split(","):
        name, B = item.split(":"); B = int(B)
        for prec in ("f64", "f32"):
            for tight in (0.1, 0.6):
                desc = synth.shape(name, 2e2 if prec == "f64" else _abi.F32_EPS_FACTOR)
                eng = ob.BatchedStackSolver(desc, B, device=0, precision=prec)
                db = torch.from_numpy(eng.pack(synth.fill(desc, 2026, 0, B, tight))).to(dev)
                dx = torch.empty((B, desc.n), dtype=torch.float64 if prec == "f64" else torch.float32, device=dev)
                ds = torch.empty(B, dtype=torch.int32, device=dev)
                di = torch.empty(B, dtype=torch.int32, device=dev)
                for _ in range(5):
                    eng.solve_device(db, dx, ds, d_iters=di)
                torch.cuda.synchronize()
                key = "%s_%d_%s_%g" % (name, B, prec, tight)
                if tight == 0.1:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(40):
                        eng.solve_device(db, dx, ds, d_iters=di)
                    e1.record(); torch.cuda.synchronize()
                    print(json.dumps({"key": key, "ms": round(e0.elapsed_time(e1) / 40, 5), "lib": os.environ.get("OSB_LIB", "head")}), flush=True)
                res[key + "_x"] = dx.cpu().numpy(); res[key + "_st"] = ds.cpu().numpy(); res[key + "_it"] = di.cpu().numpy()
                print(key, "unsolved", int((ds != 0).sum()), flush=True)
                eng.close()
    np.savez_compressed(out, **res)


if __name__ == "__main__":
    main()
