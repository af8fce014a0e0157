"""Does a page-locked buffer that the CPU has just rewritten cross PCIe slower than one that has been sitting?
(first DMA passes over freshly pinned + written buffers were measured at half the link rate: one-time mapping cost or
CPU-cache state?) One GPU."""
import ctypes as C
import glob
import os
import time

import numpy as np
import torch

torch.zeros(1, device="cuda")
paths = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = C.CDLL(paths[0])
vp = C.c_void_p
rt.cudaHostAlloc.argtypes = [C.POINTER(vp), C.c_size_t, C.c_uint]
rt.cudaMalloc.argtypes = [C.POINTER(vp), C.c_size_t]
rt.cudaMemcpyAsync.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
rt.cudaStreamSynchronize.argtypes = [vp]
size = 20 * 1000 * 1000
d = vp(); assert rt.cudaMalloc(C.byref(d), size) == 0
h = vp(); assert rt.cudaHostAlloc(C.byref(h), size, 0) == 0
arr = np.ctypeslib.as_array(C.cast(h, C.POINTER(C.c_double)), shape=(size // 8,))
src = np.random.default_rng(0).normal(size=size // 8)

def dma(tag, reps=1):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); rt.cudaMemcpyAsync(d, h, size, 1, None); rt.cudaStreamSynchronize(None); ts.append((time.perf_counter() - t0) * 1e3)
    print("%-58s %s ms" % (tag, " ".join("%.3f" % t for t in ts)), flush=True)

arr[:] = src
dma("fresh allocation, written once, 8 copies in a row:", 8)
for k in range(3):
    arr[:] = src * (k + 2.0)                      # the CPU rewrites the whole buffer (numpy: ordinary stores)
    dma("rewritten by the CPU, then 4 copies in a row:", 4)
time.sleep(0.5)
dma("after 0.5 s of idling, 3 copies:", 3)
arr[: size // 16] = 1.0
dma("half of it rewritten, 3 copies:", 3)
