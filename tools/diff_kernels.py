#!/usr/bin/env python
"""Differential run on the GPU: register-row kernel against the generic kernel (both CUDA) on
the same blobs, many seeds / shapes / saturation levels. Prints one line per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import opensot_b200 as ob
from opensot_b200 import synth


def engine(desc, B, kernel):
    if kernel == "generic":
        os.environ["OSB_KERNEL"] = "generic"
    else:
        os.environ.pop("OSB_KERNEL", None)
    e = ob.BatchedStackSolver(desc, B)
    os.environ.pop("OSB_KERNEL", None)
    return e


def main():
    B = 2048
    worst = 0.0
    for name in ("C29", "C35f", "C35h", "C38c"):
        desc = synth.shape(name)
        ea, eg = engine(desc, B, "auto"), engine(desc, B, "generic")
        assert ea.kernel_name().startswith(("rr_", "mw_")) and eg.kernel_name().startswith("sot_")
        for tight in (0.05, 0.1, 0.2, 0.4, 0.6):
            for seed in (1, 2, 3):
                blobs = synth.fill(desc, 1000 + seed, 0, B, tight)
                xa, sa = ea.solve_host(blobs)
                xg, sg = eg.solve_host(blobs)
                both = (sa == 0) & (sg == 0)
                err = float(np.abs(xa - xg)[both].max()) if both.any() else 0.0
                worst = max(worst, err)
                print("%-5s tight %.2f seed %d: rr fails %d generic fails %d status mismatches %d max|dx| %.2e"
                      % (name, tight, seed, int((sa != 0).sum()), int((sg != 0).sum()),
                         int(((sa != 0) != (sg != 0)).sum()), err), flush=True)
    # random shapes (tools/fuzz_shapes.py generator): every register-row instance, also the ones no
    # named shape uses (30 columns x 2 rows per lane)
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_shapes", os.path.join(ROOT, "tools", "fuzz_shapes.py"))
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    rng = np.random.default_rng(5)
    seen = {}
    for it in range(40):
        desc = fz.random_desc(rng, well_posed=True)
        ea, eg = engine(desc, 512, "auto"), engine(desc, 512, "generic")
        blobs = synth.fill(desc, 2000 + it, 0, 512, float(rng.choice([0.1, 0.3])))
        xa, sa = ea.solve_host(blobs)
        xg, sg = eg.solve_host(blobs)
        both = (sa == 0) & (sg == 0)
        err = float(np.abs(xa - xg)[both].max()) if both.any() else 0.0
        worst = max(worst, err)
        name = ea.kernel_name()
        seen[name] = seen.get(name, 0) + 1
        mism = int(((sa != 0) != (sg != 0)).sum())
        if mism or err > 1e-6:
            print("random shape %d (%s, n=%d): status mismatches %d, max|dx| %.2e" % (it, name, desc.n, mism, err), flush=True)
        ea.close(); eg.close()
    print("random shapes per kernel instance:", seen)
    print("worst", worst)


if __name__ == "__main__":
    main()
