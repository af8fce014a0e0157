import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import opensot_b200.engine as E
E._LIB_PATH = E._LIB_PATH.replace("libopensot_b200.so", "libopensot_b200_prof.so")
import opensot_b200 as ob
from oracle import portlib as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = P.shape("C29"); blobs = P.synth(d, 2026, 0, 4096, 0.1)[:B]
eng = ob.BatchedStackSolver(d, 4096)
for _ in range(2):
    x, st = eng.solve_host(blobs)
