import torch, numpy as np, ctypes
x = torch.empty(8*2160*3840*3, dtype=torch.float64, device="cuda")
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
nb = x.numel()*8
for name, fn in (("torch.fill_", lambda: x.fill_(1.5)), ("torch.zero_ (memset)", lambda: x.zero_())):
    ms = timed(fn); print(name, round(ms,4), "ms", round(nb/ms/1e6,0), "GB/s")
y = torch.empty_like(x)
ms = timed(lambda: y.copy_(x)); print("torch copy", round(ms,4), round(2*nb/ms/1e6,0), "GB/s (r+w)")
ms = timed(lambda: x.sum()); print("torch sum (read)", round(ms,4), round(nb/ms/1e6,0), "GB/s")
