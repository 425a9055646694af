"""Encode/decode timings of both search kernels over the SURVEY 8d shapes (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
dev = torch.device("cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


shapes = [  # (name, B, D, T, K, S, G)
    ("cfg1 recipe batch", 16, 128, 100, 1024, 8, 1),
    ("cfg1 B=256", 256, 128, 100, 1024, 8, 1),
    ("cfg1 B=4096", 4096, 128, 100, 1024, 8, 1),
    ("cfg1-recipe D512 S12 B=16", 16, 512, 100, 1024, 12, 1),
    ("cfg4 D512 S12 8x10s", 8, 512, 1000, 1024, 12, 1),
    ("cfg4 D512 S12 64x10s", 64, 512, 1000, 1024, 12, 1),
    ("cfg3 GRVQ 64x50", 64, 512, 50, 1024, 2, 2),
    ("cfg3 GRVQ 4096x50", 4096, 512, 50, 1024, 2, 2),
    ("cfg2 8x60s", 8, 512, 45000, 1024, 1, 1),
]
g = torch.Generator(device="cpu").manual_seed(1)
for name, b, d, t, k, s, gr in shapes:
    x = torch.randn(b, d, t, generator=g).to(dev)
    cbs = [(torch.randn(k, d // gr, generator=g) * (0.7 ** (i // gr))).to(dev) for i in range(s * gr)]
    hn = ops.codebook_half_norms(cbs)
    pack = ops.tc_pack_codebooks(cbs)
    n = b * t
    codes = torch.empty((s * gr, n), dtype=torch.int64, device=dev)
    out = torch.empty((b, d, t), dtype=torch.float32, device=dev)
    fl = 2.0 * k * (d // gr) * gr * s * n
    flags = ops.ACQ_STE if gr > 1 else 0
    t_tc = timeit(lambda: ops.rvq_search(x, cbs, s, gr, half_norms=hn, flags=flags, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes))
    t_si = timeit(lambda: ops.rvq_search(x, cbs, s, gr, half_norms=hn, flags=flags, impl=_lib.ACQ_IMPL_SIMT, codes_out=codes), n=3)
    t_fw = timeit(lambda: ops.rvq_search(x, cbs, s, gr, half_norms=hn, flags=flags, impl=_lib.ACQ_IMPL_SIMT, want_quantized=True, want_sqerr=True), n=3)
    t_rp = timeit(lambda: ops.rvq_replay(x, codes, cbs, s, gr, flags=flags, want_quantized=True, want_sqerr=True))
    t_de = timeit(lambda: ops.vq_decode(codes, n, 1, cbs, s, gr, b, t, check=False, out=out))
    dec_bytes = n * (8.0 * s * gr + 4.0 * d)
    print(f"{name:28s} N={n:7d}: tc {t_tc:8.4f} ms ({fl/t_tc/1e9:7.1f} TF/s, {n/t_tc/1e3:8.1f} Mfr/s) | simt enc {t_si:8.4f} ms "
          f"({fl/t_si/1e9:6.1f} TF/s) | simt fwd {t_fw:8.4f} ms | replay (fwd = tc + replay) {t_rp:7.4f} ms | decode {t_de:7.4f} ms ({dec_bytes/t_de/1e6:7.1f} GB/s)")
