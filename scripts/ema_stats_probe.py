"""One launch of the stand-alone EMA statistics kernel (acq_ema_stats) at the cfg5 throughput shape, for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, synth
dev = torch.device("cuda:0")
d, s, b, t = 512, 12, 640, 100
x = torch.from_numpy(synth.latents(b, d, t, 1234)).to(dev)
cbs = [c.contiguous() for c in torch.from_numpy(synth.rvq_codebooks(s, 1024, d, 4321, "decay")).to(dev)]
codes, _, _, _ = ops.rvq_search(x, cbs, s, flags=ops.ACQ_STE, tc_pack=ops.tc_pack_codebooks(cbs))
for _ in range(2):
    st = ops.ema_stats(x, codes, cbs, flags=ops.ACQ_STE)
torch.cuda.synchronize()
print("ok", float(st[-1024:].sum()))
