"""Development aid: first-light run of the CUDA path on the GPU box with diagnostics."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import opensot_b200 as ob
from oracle import portlib as P

print(ob.lib().osb_version(), "devices", ob.lib().osb_device_count())
for name, tight, count in [("C7c", 0.3, 64), ("C12", 0.3, 64), ("C29", 0.1, 256), ("C29", 0.4, 256), ("C35f", 0.1, 128),
                           ("C38c", 0.1, 128), ("C35h", 0.1, 128), ("C64c", 0.1, 32)]:
    d = P.shape(name)
    blobs = P.synth(d, 1, 0, count, tight)
    eng = ob.BatchedStackSolver(d, count, store_levels=True)
    t = time.time(); x, st, xl = eng.solve_host(blobs); dt = time.time() - t
    xp, xlp, stp, itp = P.port_solve(d, blobs, levels=True)
    err = np.abs(x - xp).max(axis=1)
    lerr = [np.abs(xl[:, k] - xlp[:, k]).max() for k in range(d.n_levels)]
    print("%-5s tight %.2f: status!=0 %d  max err %.2e med %.2e  per-level %s  info %s  %.1f ms"
          % (name, tight, (st != 0).sum(), np.nanmax(err), np.nanmedian(err), ["%.1e" % e for e in lerr], eng.info(), dt * 1e3))
    if (st != 0).any():
        print("   status:", np.unique(st, return_counts=True))
print("fp64 peak probe TFLOP/s:", ob.measure_fp64_tflops(0, 4096))
