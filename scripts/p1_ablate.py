"""Ablation timings of the single-product search kernel (ACQ_TC_DBG bits; results are wrong by design).
    python scripts/p1_ablate.py [SHAPE] [VARIANT:CLUSTER]
bits: 1 loaders consume the x slots without converting, 2 no operand copies into the ring, 8 no epilogue sweeps,
      32 no exact re-score (every row counts as decided)
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
SHAPES = {"cfg2": (8, 512, 45000, 1024, 1, 1), "cfg1": (4096, 128, 100, 1024, 8, 1), "cfg4": (64, 512, 1000, 1024, 12, 1),
          "cfg3": (4096, 512, 50, 1024, 2, 2), "cfg5": (640, 512, 100, 1024, 12, 1), "cfg5s": (256, 512, 100, 1024, 12, 1),
          "cfg4t100": (64, 512, 1000, 1024, 12, 1)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
var = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1:1").split(":")]
b, d, t, k, s, gr = SHAPES[name]
dev = torch.device("cuda:0")
lib = _lib.load()
lib.acq_tc_configure(var[0], var[1], 0)
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.randn(b, d, t, generator=g).to(dev)
cbs = [(torch.randn(k, d // gr, generator=g) * (0.7 ** (i // gr))).to(dev) for i in range(s * gr)]
pack = ops.tc_pack_codebooks(cbs)
codes = torch.empty((s * gr, b * t), dtype=torch.int64, device=dev)
flags = ops.ACQ_STE if (gr > 1 or os.environ.get("ABL_STE") == "1") else 0
run = lambda: ops.rvq_search(x, cbs, s, gr, flags=flags, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes)
for _ in range(30):
    run()
torch.cuda.synchronize()
for bits in ([int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 32, 8 | 32, 1 | 8 | 32, 2 | 8 | 32, 1 | 2 | 8 | 32, 0]):
    os.environ["ACQ_TC_DBG"] = str(bits)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} v{var[0]} cl{var[1]} dbg={bits:3d}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
