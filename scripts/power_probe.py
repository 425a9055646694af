"""Sustained run of the tensor-core search kernel (cfg2 shape) with nvidia-smi sampling clocks / power /
throttle reasons alongside: is the kernel's clock set by the power limit?"""
import os, subprocess, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
dev = torch.device("cuda:0")
b, d, t, k, s = 8, 512, 45000, 1024, 1
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.randn(b, d, t, generator=g).to(dev)
cbs = [torch.randn(k, d, generator=g).to(dev)]
pack = ops.tc_pack_codebooks(cbs)
codes = torch.empty((s, b * t), dtype=torch.int64, device=dev)
f = lambda: ops.rvq_search(x, cbs, s, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes)
for _ in range(3): f()
torch.cuda.synchronize()
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                        "--format=csv,noheader", "-lms", "200"], stdout=subprocess.PIPE, text=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
marks = []
for chunk in range(4):
    e0.record()
    for _ in range(n // 4): f()
    e1.record(); torch.cuda.synchronize()
    marks.append(e0.elapsed_time(e1) / (n // 4))
time.sleep(0.3)
mon.terminate()
out = mon.stdout.read().strip().splitlines()
print("ms per launch over four consecutive quarters:", ["%.4f" % m for m in marks])
for line in out[:: max(1, len(out) // 16)]:
    print("  ", line)
