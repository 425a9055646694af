"""Reference points for the decode roofline: pure-write and copy bandwidth on this GPU (torch kernels)."""
import torch
dev = torch.device("cuda:0")
n = 8 * 512 * 45000
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ms = t(lambda: a.fill_(1.5)); print(f"fill 737 MB: {ms:.4f} ms = {n*4/ms/1e6:.0f} GB/s written")
ms = t(lambda: a.zero_()); print(f"memset 737 MB: {ms:.4f} ms = {n*4/ms/1e6:.0f} GB/s written")
ms = t(lambda: b.copy_(a)); print(f"copy 737 MB: {ms:.4f} ms = {2*n*4/ms/1e6:.0f} GB/s read+write")
