"""Where is the tensor-core kernel faster than the SIMT kernel for small batches? (sets the AUTO threshold)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
dev = torch.device("cuda:0")
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
g = torch.Generator(device="cpu").manual_seed(1)
for (d, s, gr) in [(128, 8, 1), (512, 12, 1), (512, 1, 1), (512, 2, 2)]:
    cbs = [(torch.randn(1024, d // gr, generator=g) * 0.7 ** (i // gr)).to(dev) for i in range(s * gr)]
    hn = ops.codebook_half_norms(cbs); pack = ops.tc_pack_codebooks(cbs)
    for n in (4, 16, 48, 100, 256, 512, 1024):
        x = torch.randn(1, d, n, generator=g).to(dev)
        codes = torch.empty((s * gr, n), dtype=torch.int64, device=dev)
        fl = ops.ACQ_STE if gr > 1 else 0
        t_tc = timeit(lambda: ops.rvq_search(x, cbs, s, gr, half_norms=hn, flags=fl, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes))
        t_si = timeit(lambda: ops.rvq_search(x, cbs, s, gr, half_norms=hn, flags=fl, impl=_lib.ACQ_IMPL_SIMT, codes_out=codes))
        print(f"D={d} S={s} G={gr} N={n:5d}: tc {t_tc:7.1f} us  simt {t_si:7.1f} us")
