"""Development aid (not a test): per-phase cycle stamps of one warp of the register-row kernel.
Build the instrumented library here with `python opensot_b200/build.py --profile`, then on the GPU box
`python tests/gpu_prof.py [batch]` prints, for block 100, the cycles of every phase of every level
(project, sweep phase 1/2, Gram, g_z, Cholesky, leading columns, active-set loop and its parts)."""
import sys
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
