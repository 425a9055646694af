"""Small invocation of every kernel (for compute-sanitizer memcheck / racecheck on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib, synth
dev = torch.device("cuda:0")
for (b, d, t, k, s, g) in [(3, 128, 100, 256, 3, 1), (2, 512, 37, 256, 2, 2), (5, 64, 130, 512, 2, 1)]:
    x = torch.from_numpy(synth.latents(b, d, t, 1)).to(dev)
    cbs = [torch.from_numpy(synth.normal((k, d // g), 10 + i)).to(dev) for i in range(s * g)]
    hn = ops.codebook_half_norms(cbs)
    pack = ops.tc_pack_codebooks(cbs)
    flags = ops.ACQ_STE if g > 1 else 0
    c_tc, _, _, _ = ops.rvq_search(x, cbs, s, g, half_norms=hn, flags=flags, impl=_lib.ACQ_IMPL_TC, tc_pack=pack)
    c_si, q, r, se = ops.rvq_search(x, cbs, s, g, half_norms=hn, flags=flags, impl=_lib.ACQ_IMPL_SIMT,
                                    want_quantized=True, want_residual=True, want_sqerr=True)
    q2, r2, se2, st2 = ops.rvq_replay(x, c_si, cbs, s, g, flags=flags, want_residual=True, want_sqerr=True,
                                      want_stats=(g == 1))
    dec = ops.vq_decode(c_si, b * t, 1, cbs, s, g, b, t)
    if g == 1:
        st = ops.ema_stats(x, c_si, cbs, flags=flags)
        ops.ema_apply(st, [c.clone() for c in cbs], [c.clone() for c in cbs],
                      [torch.zeros(k, device=dev) for _ in cbs], 0.99, 1e-5)
    torch.cuda.synchronize()
    print(f"shape B={b} D={d} T={t} K={k} S={s} G={g}: tc==simt {bool((c_tc == c_si).all())}, replay==fused {bool(torch.equal(q, q2))}")
print("SANITY DONE")
