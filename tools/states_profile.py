#!/usr/bin/env python
"""State-input path a few times (for an ncu launch list: which kernel takes what), COMAN tree fixture, B = 4096."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import opensot_b200 as ob
from opensot_b200 import synth
import bench

B = int(os.environ.get("STATES_BATCH", "4096"))
desc = synth.shape("C29")
eng = ob.BatchedStackSolver(desc, B)
r = bench.states_e2e(eng, B, desc.n, int(os.environ.get("STATES_STEPS", "20")), 0, 1, None)
print(r)
