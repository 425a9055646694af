"""Search-kernel timings per variant over the BASELINE shapes, with the kernels' own stall counters.

    python scripts/search_sweep.py [variants ...]      e.g.  1:1 1:2 3:1   (variant:cluster, default all)

Per shape and variant: ms per launch, algorithmic TFLOP/s (2*K*Dg*G*S per frame), and -- from one extra
launch with ACQ_TC_DBG=512 -- per-CTA kilocycles of the MMA thread's waits, the TMA thread's waits, the
epilogue's wait / sweep / slot time and the fraction of (frame, stage) decisions that needed the exact
re-score.
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import ops, _lib
dev = torch.device("cuda:0")
lib = _lib.load()


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
    ("cfg2 8x60s D512 S1", 8, 512, 45000, 1024, 1, 1),
    ("cfg1 B=4096 D128 S8", 4096, 128, 100, 1024, 8, 1),
    ("cfg4 64x10s D512 S12", 64, 512, 1000, 1024, 12, 1),
    ("cfg3 GRVQ 4096x50", 4096, 512, 50, 1024, 2, 2),
    ("cfg1 B=256 D128 S8", 256, 128, 100, 1024, 8, 1),
    ("cfg4 8x10s D512 S12", 8, 512, 1000, 1024, 12, 1),
]
if os.environ.get("SWEEP_SHAPES"):          # e.g. SWEEP_SHAPES=cfg2,cfg4 : only the shapes whose name starts with one of these
    shapes = [sh for sh in shapes if any(sh[0].startswith(p) for p in os.environ["SWEEP_SHAPES"].split(","))]
variants = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(1, 1), (1, 2), (1, 4), (3, 1)]
g = torch.Generator(device="cpu").manual_seed(1)
for name, b, d, t, k, s, gr in shapes:
    x = torch.randn(b, d, t, generator=g).to(dev)
    cbs = [(torch.randn(k, d // gr, generator=g) * (0.7 ** (i // gr))).to(dev) for i in range(s * gr)]
    pack = ops.tc_pack_codebooks(cbs)
    n = b * t
    codes = torch.empty((s * gr, n), dtype=torch.int64, device=dev)
    fl = 2.0 * k * (d // gr) * gr * s * n
    flags = ops.ACQ_STE if gr > 1 else 0
    run = lambda: ops.rvq_search(x, cbs, s, gr, flags=flags, impl=_lib.ACQ_IMPL_TC, tc_pack=pack, codes_out=codes)
    ref = None
    lib.acq_tc_configure(3, 1, 0)
    for _ in range(40):          # bring the clocks up before anything is timed
        run()
    torch.cuda.synchronize()
    for (var, cl) in variants:
        lib.acq_tc_configure(var, cl, 0)
        ms = timeit(run, n=20)
        got = codes.clone()
        if ref is None:
            ref = got
        ndiff = int((got != ref).any(0).sum())
        # stall counters of one launch
        ws = ops.tc_workspace(d, dev)
        base = int(lib.acq_tc_workspace_bytes(d)) - 256 + 64
        prev_dbg = os.environ.get("ACQ_TC_DBG")
        os.environ["ACQ_TC_DBG"] = str(512 | int(prev_dbg or 0))
        ws[base:base + 192].zero_()
        run()
        torch.cuda.synchronize()
        if prev_dbg is None:
            os.environ.pop("ACQ_TC_DBG")
        else:
            os.environ["ACQ_TC_DBG"] = prev_dbg
        st = ws[base:base + 192].view(torch.int64).tolist()
        ctas = min(148, (n + 127) // 128)
        kc = lambda v: v / ctas / 1e3
        extra = ""
        if var == 1:
            extra = (f" epi wait/sweep/slot {kc(st[10]):.0f}/{kc(st[11]):.0f}/{kc(st[12]):.0f}"
                     f" rescore {100.0 * st[13] / (n * s * gr):.2f}% full {st[14]}"
                     f" | loader warps sweep0/sweep1 " + " ".join(f"{kc(st[15 + 2 * w]):.0f}/{kc(st[16 + 2 * w]):.0f}" for w in range(4))
                     + f" bar {kc(st[23]):.0f}"
                     + (f" | epi set 0: exchange barrier {kc(st[15]):.0f}, filter+publish {kc(st[16]):.0f}, first image {kc(st[17]):.0f}" if s * gr > 1 else "")
                     + (f" | jobs: {st[1] / ctas:.0f} per CTA, publish->done {st[0] / max(st[1], 1) / 1e3:.1f} kcyc each, "
                        f"{st[4] / max(st[1], 1):.1f} batches of {st[3] / max(st[4], 1) / 1e3:.2f} kcyc (decide {st[18] / max(st[4], 1) / 1e3:.2f} update {st[19] / max(st[4], 1) / 1e3:.2f} fence {st[20] / max(st[4], 1) / 1e3:.2f})" if st[1] else ""))
        print(f"{name:22s} v{var} cl{cl}: {ms:7.4f} ms {fl / ms / 1e9:7.1f} TF/s ~{kc(st[7]) / ms / 1e3:.2f} GHz diff_vs_first={ndiff:3d} | kcyc/CTA: mma total {kc(st[7]):.0f} "
              f"wait full0/full/tempty {kc(st[0]):.0f}/{kc(st[1]):.0f}/{kc(st[2]):.0f} tma wait empty/img {kc(st[3]):.0f}/{kc(st[4]):.0f} "
              f"loader wait free/x {kc(st[5]):.0f}/{kc(st[6]):.0f} of {kc(st[8]):.0f} streamer wait {kc(st[9]):.0f}{extra}", flush=True)
lib.acq_tc_configure(*_lib.tc_config_defaults())
