#!/usr/bin/env python
"""Achieved HBM GB/s of the two HBM-bound helper kernels of the path (north_star: achieved GB/s per
choice): sot_gather_kernel (device-side piling, Aggregated/AutoStack::update) and sot_posture_kernel
(Postural / JointLimits / VelocityLimits from joint states). Algorithmic bytes = bytes read from the
sources + bytes written into the blobs; inputs larger than L2 (B = 65536). One JSON line per kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import opensot_b200 as ob
    from opensot_b200 import synth
    B = int(os.environ.get("AUX_BATCH", "65536"))
    steps = 20
    dev = torch.device("cuda", 0)
    peak = 6462.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    desc = synth.shape("C29")
    L = ob.blob_layout(desc)
    n = desc.n
    eng = ob.BatchedStackSolver(desc, B, device=0)
    blobs = torch.zeros((B, L.stride), dtype=torch.float64, device=dev)
    A0 = torch.randn((B, 3, n), dtype=torch.float64, device=dev)
    A1a = torch.randn((B, 6, n), dtype=torch.float64, device=dev)
    A1b = torch.randn((B, 6, n), dtype=torch.float64, device=dev)
    b0 = torch.randn((B, 3, 1), dtype=torch.float64, device=dev)
    b1a = torch.randn((B, 6, 1), dtype=torch.float64, device=dev)
    b1b = torch.randn((B, 6, 1), dtype=torch.float64, device=dev)
    srcs = [dict(src=A0, rows=3, cols=n, dst_off=L.A[0]), dict(src=A1a, rows=6, cols=n, dst_off=L.A[1]),
            dict(src=A1b, rows=6, cols=n, dst_off=L.A[1] + 6 * n), dict(src=b0, rows=3, cols=1, dst_off=L.b[0]),
            dict(src=b1a, rows=6, cols=1, dst_off=L.b[1]), dict(src=b1b, rows=6, cols=1, dst_off=L.b[1] + 6)]
    moved = sum(s["rows"] * s["cols"] for s in srcs) * 8 * 2 * B

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    ms = timeit(lambda: eng.gather_device(srcs, blobs))
    print(json.dumps({"kernel": "sot_gather_kernel", "batch": B, "sources": len(srcs), "ms": round(ms, 4),
                      "algorithmic_bytes": moved, "achieved_gbs": round(moved / ms / 1e6, 1), "peak_gbs": peak,
                      "frac": round(moved / ms / 1e6 / peak, 3)}), flush=True)
    q = torch.randn((B, n), dtype=torch.float64, device=dev)
    qr = torch.randn((B, n), dtype=torch.float64, device=dev)
    lim = torch.ones(n, dtype=torch.float64, device=dev)
    moved = B * n * 8 * (2 + 5)        # q, q_ref read; b, jl, ju, vl, vu written
    ms = timeit(lambda: eng.produce_posture_device(blobs, q, q_ref=qr, q_min=-lim, q_max=lim, qdot_max=lim, lam=1.0, dT=0.01,
                                                   postural_level=2, joint_limit_set=0, velocity_limit_set=1))
    print(json.dumps({"kernel": "sot_posture_kernel", "batch": B, "ms": round(ms, 4), "algorithmic_bytes": moved,
                      "achieved_gbs": round(moved / ms / 1e6, 1), "peak_gbs": peak, "frac": round(moved / ms / 1e6 / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
