"""H2D rate of page-locked host memory, default against write-combined (cudaHostAllocWriteCombined), one GPU or (under
torchrun) every rank at once: what the host-buffer path of a multi-GPU box can get out of its PCIe fabric."""
import ctypes as C
import glob
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.zeros(1, device="cuda")
paths = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = C.CDLL(paths[0])
vp = C.c_void_p
rt.cudaHostAlloc.argtypes = [C.POINTER(vp), C.c_size_t, C.c_uint]
rt.cudaMalloc.argtypes = [C.POINTER(vp), C.c_size_t]
rt.cudaMemcpyAsync.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
rt.cudaStreamSynchronize.argtypes = [vp]
MB = 20
size = MB * 1000 * 1000
d = vp(); assert rt.cudaMalloc(C.byref(d), size) == 0
for name, flag in (("default", 0), ("write-combined", 4), ("portable", 1)):
    h = vp(); assert rt.cudaHostAlloc(C.byref(h), size, flag) == 0
    C.memset(h, 1, size)
    for _ in range(3):
        rt.cudaMemcpyAsync(d, h, size, 1, None)
    rt.cudaStreamSynchronize(None)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        rt.cudaMemcpyAsync(d, h, size, 1, None)
    rt.cudaStreamSynchronize(None)
    dt = (time.perf_counter() - t0) / 50
    rate = torch.tensor([size / dt / 1e9], device="cuda")
    if world > 1:
        lo = rate.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(rate, op=dist.ReduceOp.SUM)
        if rank == 0:
            print("%-15s H2D %d MB x %d ranks at once: %.1f GB/s in total, slowest rank %.1f GB/s" % (name, MB, world, rate.item(), lo.item()), flush=True)
    else:
        print("%-15s H2D %d MB: %.1f GB/s" % (name, MB, rate.item()), flush=True)
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -30; lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'")
if world > 1:
    dist.destroy_process_group()
