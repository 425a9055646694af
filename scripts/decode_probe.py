import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
for (b, d, t, k, s) in [(8, 512, 45000, 1024, 1), (4096, 128, 100, 1024, 8), (64, 512, 1000, 1024, 12)]:
    cbs = [torch.randn(k, d, generator=g).to(dev) for _ in range(s)]
    codes = torch.randint(0, k, (s, b * t), generator=g).to(dev)
    out = torch.empty((b, d, t), dtype=torch.float32, device=dev)
    f = lambda: ops.vq_decode(codes, b * t, 1, cbs, s, 1, b, t, check=False, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = b * t * (8.0 * s + 4.0 * d)
    print(f"decode B={b} D={d} T={t} S={s}: {ms:.4f} ms  {by/ms/1e6:.0f} GB/s algorithmic")
