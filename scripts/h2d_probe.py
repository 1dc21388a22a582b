import torch, time
n = (1<<20)*7
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device='cuda')
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): d.copy_(h, non_blocking=True)
    s.synchronize()
    for sz in (n, n//2, n//8, n//32):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): d[:sz].copy_(h[:sz], non_blocking=True)
        e1.record(s); s.synchronize()
        ms=e0.elapsed_time(e1)/20
        print(sz*8/1e6, 'MB', ms, 'ms', sz*8/ms/1e6, 'GB/s')
