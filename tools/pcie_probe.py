import torch, time
for mb in (5, 20, 80):
    h = torch.empty(mb*1024*1024//8, dtype=torch.float64).pin_memory(); d = torch.empty_like(h, device='cuda')
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print("H2D %d MB: %.3f ms  %.1f GB/s" % (mb, ms, mb*1.048576/ms))
    h2 = torch.empty_like(h).pin_memory()
    e0.record()
    for _ in range(20): h2.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print("D2H %d MB: %.3f ms  %.1f GB/s" % (mb, ms, mb*1.048576/ms))
