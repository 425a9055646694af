"""Time the tensor-core search kernel alone (cfg2 shape) under the ACQ_TC_DBG experiment bits."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
dev = torch.device("cuda:0")
b, d, t, k, s = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (8, 512, 45000, 1024, 1)))
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.randn(b, d, t, generator=g).to(dev)
cbs = [torch.randn(k, d, generator=g).to(dev) * (0.7 ** i) for i in range(s)]
pack = ops.tc_pack_codebooks(cbs)
codes = torch.empty((s, b * t), dtype=torch.int64, device=dev)
for _ in range(3):
    ops.rvq_search(x, cbs, s, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    ops.rvq_search(x, cbs, s, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 2.0 * k * d * s * b * t
print(f"ACQ_TC_DBG={os.environ.get('ACQ_TC_DBG','0')} shape=({b},{d},{t}) K={k} S={s}: {ms:.4f} ms  {fl/ms/1e9:.1f} TFLOP/s algorithmic")
if int(os.environ.get("ACQ_TC_DBG", "0")) & 512:
    ws = ops.tc_workspace(d, dev)
    base = int(_lib.load().acq_tc_workspace_bytes(d)) - 256 + 64
    c = ws[base:base + 104].view(torch.int64).cpu().tolist()
    runs = n + 3
    names = ["mma wait full (pass 0)", "mma wait full (pass 1+)", "mma wait tempty", "tma wait empty", "tma wait t0",
             "loader wait free", "-", "mma thread total", "loader total", "epilogue: norms staging",
             "epilogue: wait accumulator", "epilogue: sweeps", "epilogue: residual update"]
    for nm, v in zip(names, c):
        print(f"  {nm:26s} {v / runs / 148 / 1e3:10.1f} kcycles per CTA per launch")
