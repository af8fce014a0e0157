#!/usr/bin/env python
"""Development aid: blocking osb_solve_host against the chunk pattern of its copy / solve / copy-out pipeline
(OSB_PIPE_PATTERN), C29 x 4096 fp64, page-locked buffers. Host wall clock per call."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import opensot_b200 as ob
from opensot_b200 import synth

B = 4096
desc = synth.shape("C29")
eng = ob.BatchedStackSolver(desc, B)
sets = [torch.from_numpy(eng.pack(synth.fill(desc, 5, s * B, B, 0.1))).pin_memory() for s in range(9)]
x = torch.empty((B, desc.n), dtype=torch.float64).pin_memory(); st = torch.empty(B, dtype=torch.int32).pin_memory()
pats = ["", "4096", "2048,2048", "3072,1024", "2560,1024,512", "2048,1024,1024", "2048,1024,512,512", "1024,1024,1024,1024",
        "1536,1024,768,512,256", "2048,1024,512,256,256", "3072,512,512", "2816,768,512", "3328,512,256"]
for p in pats:
    if p:
        os.environ["OSB_PIPE_PATTERN"] = p
    else:
        os.environ.pop("OSB_PIPE_PATTERN", None)
    for i in range(5):
        eng.solve_host(sets[i % 9].numpy(), x.numpy(), st.numpy())
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(40):
            eng.solve_host(sets[(i + rep) % 9].numpy(), x.numpy(), st.numpy())
        best = min(best, (time.perf_counter() - t0) / 40 * 1e3)
    print("%-28s %.4f ms per call" % (p or "(default)", best), flush=True)
