import torch, time
dev = torch.device("cuda")
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
for mb in (134, 1024):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
    t = timeit(lambda: a.zero_()); print(f"{mb} MB write-only (zero_):   {n*4/t/1e9:7.0f} GB/s  {t*1e6:7.1f} us")
    t = timeit(lambda: b.copy_(a)); print(f"{mb} MB copy (read+write):   {2*n*4/t/1e9:7.0f} GB/s  {t*1e6:7.1f} us")
    t = timeit(lambda: a.sum()); print(f"{mb} MB read-only (sum):     {n*4/t/1e9:7.0f} GB/s  {t*1e6:7.1f} us")
    c = torch.empty(n // 4, device=dev, dtype=torch.int32)
    t = timeit(lambda: torch.add(a, 1.0, out=b)); print(f"{mb} MB add (read+write):    {2*n*4/t/1e9:7.0f} GB/s  {t*1e6:7.1f} us")
