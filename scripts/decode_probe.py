"""Time and check vq_decode on the bench shapes; ACQ_DECODE_KERNEL=1 forces the tile kernel (K2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
for (b, d, t, k, s) in [(8, 512, 45000, 1024, 1), (4096, 128, 100, 1024, 8), (64, 512, 1000, 1024, 12),
                        (3, 512, 1001, 1024, 2), (8, 512, 45000, 1024, 2), (8, 512, 45000, 1024, 3),
                        (16, 128, 48000, 1024, 4), (1, 512, 45000, 1024, 1), (1, 512, 7504, 1024, 1)]:
    cbs = [torch.randn(k, d, generator=g).to(dev) for _ in range(s)]
    codes = torch.randint(0, k, (s, b * t), generator=g).to(dev)
    out = torch.empty((b, d, t), dtype=torch.float32, device=dev)
    f = lambda: ops.vq_decode(codes, b * t, 1, cbs, s, 1, b, t, check=False, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    want = torch.zeros(b * t, d, device=dev)
    for i in range(s):
        want = want + cbs[i][codes[i]]
    ok = torch.equal(out, want.view(b, t, d).permute(0, 2, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = b * t * (8.0 * s + 4.0 * d)
    print(f"decode B={b} D={d} T={t} S={s}: {ms:.4f} ms  {by/ms/1e6:.0f} GB/s algorithmic  exact={ok}")

# GRVQ embed layout: codes [B, T, 2G], tables ordered stage-major (cfg3)
b, t, k, dg, G, S = 4096, 48, 1024, 256, 2, 2
cbs = [torch.randn(k, dg, generator=g).to(dev) for _ in range(S * G)]
codes = torch.randint(0, k, (b, t, S * G), generator=g).to(dev)
out = torch.empty((b, dg * G, t), dtype=torch.float32, device=dev)
f = lambda: ops.vq_decode(codes, 1, S * G, cbs, S, G, b, t, check=False, out=out)
for _ in range(3): f()
torch.cuda.synchronize()
want = torch.zeros(b * t, dg * G, device=dev)
cf = codes.view(-1, S * G)
for s_ in range(S):
    want = want + torch.cat([cbs[s_ * G + g_][cf[:, s_ * G + g_]] for g_ in range(G)], dim=1)
ok = torch.equal(out, want.view(b, t, dg * G).permute(0, 2, 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
by = b * t * (8.0 * S * G + 4.0 * dg * G)
print(f"grvq embed B={b} D={dg*G} T={t} S={S} G={G}: {ms:.4f} ms  {by/ms/1e6:.0f} GB/s algorithmic  exact={ok}")
