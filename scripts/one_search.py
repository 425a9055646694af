"""One shape, one kernel variant, a few launches: the target of ncu captures.

    python scripts/one_search.py SHAPE VARIANT[:CLUSTER] [launches]
    SHAPE: cfg2 | cfg1 | cfg1s | cfg4 | cfg4s | cfg3
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
SHAPES = {  # B, D, T, K, S, G
    "cfg2": (8, 512, 45000, 1024, 1, 1), "cfg1": (4096, 128, 100, 1024, 8, 1), "cfg1s": (256, 128, 100, 1024, 8, 1),
    "cfg4": (64, 512, 1000, 1024, 12, 1), "cfg4s": (8, 512, 1000, 1024, 12, 1), "cfg3": (4096, 512, 50, 1024, 2, 2),
}
b, d, t, k, s, gr = SHAPES[sys.argv[1]]
var = [int(v) for v in sys.argv[2].split(":")] + [1]
n_launch = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
_lib.load().acq_tc_configure(var[0], var[1], 0)
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.randn(b, d, t, generator=g).to(dev)
cbs = [(torch.randn(k, d // gr, generator=g) * (0.7 ** (i // gr))).to(dev) for i in range(s * gr)]
pack = ops.tc_pack_codebooks(cbs)
codes = torch.empty((s * gr, b * t), dtype=torch.int64, device=dev)
flags = ops.ACQ_STE if gr > 1 else 0
for _ in range(n_launch):
    ops.rvq_search(x, cbs, s, gr, flags=flags, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes)
torch.cuda.synchronize()
print("ok", sys.argv[1:], int(codes.sum()))
ws = ops.tc_workspace(d, dev)
base = int(_lib.load().acq_tc_workspace_bytes(d)) - 256
print("err flag", ws[base:base + 4].view(torch.int32).item())
