"""Read-only / write-only / copy HBM bandwidth with torch library kernels (context for the roofline)."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for mb in (88, 352, 1024, 4096):
    n = mb * 1024 * 1024 // 4
    x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
    xs = [torch.empty(n, device="cuda") for _ in range(4)] if mb <= 352 else [x]
    i = [0]
    def fill():
        xs[i[0] % len(xs)].fill_(1.0); i[0] += 1
    def rd():
        xs[i[0] % len(xs)].sum(); i[0] += 1
    w = t(fill); r = t(rd); c = t(lambda: y.copy_(x))
    print(f"{mb:5d} MB  write-only {n*4/w/1e9:7.0f} GB/s ({w*1e6:6.1f} us)   read-only {n*4/r/1e9:7.0f} GB/s   copy {2*n*4/c/1e9:7.0f} GB/s (r+w)")
