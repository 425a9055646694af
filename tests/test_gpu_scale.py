"""BASELINE-scale parity: every frame of every BASELINE.json shape against the oracle.

VERDICT r01 "weak #1/#2": the advertised shapes (cfg2 8 x 45 000; cfg1 4096 x 100, D=128, S=8; cfg4
64 x 1000, D=512, S=12; cfg3 4096 x 50, G=2, straight-through) were only sample-audited, and the
persistent multi-tile tensor-core path at D_g = 256 / 512 was never compared with the oracle.  Here the
CUDA codes of ALL frames are compared index for index with the oracle's (the reference's ATen call
sequence on the host CPU, oracle/rvq_oracle.py, oracle/grvq_oracle.py); every disagreement must be an
fp64 near-tie within adjudicate.EPS_ULPS fp32 ulps and is counted and printed; a frame sample is
audited in fp64 along its own residual chain; decode of the oracle's codes is bit-exact.

Each shape runs on every tensor-core variant (acq_tc_configure): single-product filter + exact re-score
(variant 1) and the three-product split (variant 3), each alone and with a 2-CTA multicast cluster.
The small-batch split mode, determinism and TC-vs-SIMT agreement on random shapes (formerly
scripts/tc_stress.py and scripts/split_stress.py) are folded in at the bottom.
"""
import functools
import random

import numpy as np
import pytest
import torch

from academicodec_b200 import synth

pytestmark = pytest.mark.gpu

# (variant, cluster): see include/acq_b200.h acq_tc_configure
VARIANTS = [(1, 1), (1, 2), (3, 1), (3, 2)]
VARIANT_IDS = ["p1", "p1_cl2", "p3", "p3_cl2"]

RVQ_SCALE = {
    # name: (B, D, T, n_q, regime)
    "cfg2_8x45000": (8, 512, 45000, 1, "decay"),
    "cfg1_4096x100": (4096, 128, 100, 8, "decay"),
    "cfg4_64x1000": (64, 512, 1000, 12, "decay"),
    "cfg1recipe_256x100": (256, 512, 100, 12, "randn"),
}
K = 1024


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture
def configure():
    """acq_tc_configure for one test; the library's defaults are restored afterwards."""
    from academicodec_b200 import _lib
    lib = _lib.load()
    default = _lib.tc_config_defaults()

    def set_(variant, cluster, split=-1):
        _lib.check(lib.acq_tc_configure(variant, cluster, split), "acq_tc_configure")
    yield set_
    lib.acq_tc_configure(*default)


@functools.lru_cache(maxsize=2)
def rvq_case(name):
    """Inputs and the oracle's codes (torch CPU, all host threads), computed once per shape."""
    from oracle import rvq_oracle
    b, d, t, s, regime = RVQ_SCALE[name]
    x = torch.from_numpy(synth.latents(b, d, t, 1234))
    cb = torch.from_numpy(synth.rvq_codebooks(s, K, d, 4321, regime))
    with torch.no_grad():
        ref = rvq_oracle.rvq_encode(x, list(cb))
    return x, cb, ref


def report(tag, rep):
    print(f"[scale {tag}] codes={rep['total']} identical={rep['identical']} near_tie={rep['near_tie']} "
          f"downstream={rep['downstream']} hard={rep['hard_mismatch']} diverged_frames={rep['diverged_frames']}")


@pytest.mark.parametrize("variant", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("name", list(RVQ_SCALE))
def test_rvq_all_frames_vs_oracle(dev, configure, name, variant):
    from academicodec_b200 import _lib, ops
    from oracle import adjudicate, rvq_oracle
    b, d, t, s, _ = RVQ_SCALE[name]
    x, cb, ref = rvq_case(name)
    configure(variant[0], variant[1], 0)             # persistent kernel (no small-batch split)
    cbs = [cb[i].to(dev).contiguous() for i in range(s)]
    pack = ops.tc_pack_codebooks(cbs)
    xd = x.to(dev)
    codes, _, _, _ = ops.rvq_search(xd, cbs, s, impl=_lib.ACQ_IMPL_TC, tc_pack=pack)
    again, _, _, _ = ops.rvq_search(xd, cbs, s, impl=_lib.ACQ_IMPL_TC, tc_pack=pack)
    assert torch.equal(codes, again), "non-deterministic codes"
    got = codes.view(s, b, t).cpu().numpy()
    rep = adjudicate.compare_rvq_codes(x, cb, ref.numpy(), got)
    report(f"{name}/{variant}", rep)
    assert rep["hard_mismatch"] == 0, rep["hard_examples"]
    assert rep["diverged_frames"] <= max(1, b * t // 2000), rep
    # fp64 audit of a clip sample along the CUDA path's own residual chain
    pick = np.random.RandomState(5).choice(b, size=min(b, max(1, 4000 // t)), replace=False)
    if t > 4000:
        sl = slice(1000, 5000)
        audit = adjudicate.audit_rvq_codes(x[pick][:, :, sl], cb, got[:, pick][:, :, sl])
    else:
        audit = adjudicate.audit_rvq_codes(x[pick], cb, got[:, pick])
    assert sum(audit["wrong"]) == 0, audit
    # decode of the oracle's codes: bit-exact, all frames
    dec = ops.vq_decode(ref.to(dev), b * t, 1, cbs, s, 1, b, t)
    want = rvq_oracle.rvq_decode(ref, list(cb))
    assert torch.equal(dec.cpu(), want)


@functools.lru_cache(maxsize=1)
def grvq_case():
    from oracle import grvq_oracle
    b, t, g = 4096, 50, 2
    x = torch.from_numpy(synth.latents(b, 512, t, 31))
    w = synth.grvq_codebooks(g, K, 777, "randn")
    w = [[torch.from_numpy(a) for a in st] for st in w]
    with torch.no_grad():
        q, loss, ids = grvq_oracle.grvq_forward(x, w)
    return x, w, q, loss, torch.stack(ids)        # ids: [s0g0, s0g1, s1g0, s1g1] x [B*T]


@pytest.mark.parametrize("variant", VARIANTS, ids=VARIANT_IDS)
def test_grvq_cfg3_all_frames_vs_oracle(dev, configure, variant):
    """cfg3: HiFi-Codec GRVQ, 4096 clips x 50 frames (T % 4 != 0: the loaders' scalar path), 2 groups x 2
    stages, straight-through residual arithmetic (models.py:478,490)."""
    from academicodec_b200 import _lib, ops
    from oracle import adjudicate, grvq_oracle
    x, w, q_ref, loss_ref, ref = grvq_case()
    b, _, t = x.shape
    g = 2
    configure(variant[0], variant[1], 0)
    ws = [a.to(dev) for st in w for a in st]
    pack = ops.tc_pack_codebooks(ws)
    xd = x.to(dev)
    codes, _, _, _ = ops.rvq_search(xd, ws, 2, g, flags=ops.ACQ_STE | ops.ACQ_LOSS_RAW,
                                    impl=_lib.ACQ_IMPL_TC, tc_pack=pack)
    got = codes.cpu().numpy()
    refn = ref.numpy()
    diverged = 0
    for gi in range(g):                                # each group is a 2-stage RVQ on its channels
        xs = x[:, gi * 256:(gi + 1) * 256]
        cbs = [w[0][gi], w[1][gi]]
        r = np.stack([refn[gi], refn[g + gi]]).reshape(2, b, t)
        n = np.stack([got[gi], got[g + gi]]).reshape(2, b, t)
        rep = adjudicate.compare_rvq_codes(xs, cbs, r, n, straight_through=True)
        report(f"cfg3 group {gi}/{variant}", rep)
        assert rep["hard_mismatch"] == 0, rep["hard_examples"]
        diverged += rep["diverged_frames"]
    assert diverged <= max(1, b * t // 2000)
    # forward outputs through the AUTO path (tensor-core codes + replay) on frames with identical codes
    c2, quant, _, sqerr = ops.rvq_search(xd, ws, 2, g, flags=ops.ACQ_STE | ops.ACQ_LOSS_RAW, tc_pack=pack,
                                         want_quantized=True, want_sqerr=True)
    assert torch.equal(c2, codes)
    same = (ref == codes.cpu()).all(0).view(b, t)
    assert np.array_equal(quant.cpu().numpy().transpose(0, 2, 1)[same.numpy()],
                          q_ref.numpy().transpose(0, 2, 1)[same.numpy()])
    if bool(same.all()):
        loss = float((sqerr * (1.25 / x.numel())).mean())
        np.testing.assert_allclose(loss, float(loss_ref), rtol=1e-5)
    # embed (decode) of the oracle's codes in the [B, T, 4] layout: bit-exact
    inter = ref.t().reshape(b, t, 4).contiguous()
    emb = ops.vq_decode(inter.to(dev), 1, 4, ws, 2, g, b, t)
    assert torch.equal(emb.cpu(), grvq_oracle.grvq_embed(inter, w))


# ---- small-batch split mode, determinism, TC vs SIMT on random shapes ----------------------------
def _stress_shapes(n, seed, small):
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        g_ = rnd.choice([1, 1, 1, 2, 4])
        dg = rnd.choice([64, 128, 256, 512])
        if dg * g_ > 1024:
            dg = 1024 // g_ // 64 * 64
        k = rnd.choice([256, 512, 1024, 1024])
        s = rnd.randint(1, 12 if dg * g_ <= 512 else 4)
        if small:
            n_fr = rnd.randint(1, 74 * 128)            # 1..74 tiles: the cluster-split kernels
            t = rnd.choice([1, 3, 50, 100, 127, 750])
        else:
            n_fr = rnd.randint(75 * 128, 600 * 128)    # persistent kernel, several tiles per CTA
            t = rnd.choice([37, 100, 128, 129, 1000, 4099])
        b = max(1, n_fr // t)
        if b * t * dg * g_ > 3e8:
            continue
        out.append((b, t, dg, g_, k, s))
    return out


@pytest.mark.parametrize("variant", [(1, 1, 1), (3, 1, 1), (1, 2, 0), (3, 2, 0), (3, 4, 0)],
                         ids=["p1_split", "p3_split", "p1_cl2", "p3_cl2", "p3_cl4"])
@pytest.mark.parametrize("small", [True, False], ids=["small", "multi_tile"])
def test_stress_tc_vs_simt(dev, configure, variant, small):
    """Random shapes: tensor-core codes are deterministic over repeated launches and differ from the fused
    SIMT kernel's only on a handful of frames (near-ties and their downstream stages)."""
    from academicodec_b200 import _lib, ops
    configure(*variant)
    gen = torch.Generator(device="cpu").manual_seed(7)
    tot = diff = 0
    for (b, t, dg, g_, k, s) in _stress_shapes(24 if small else 8, 11 if small else 12, small):
        x = torch.randn(b, dg * g_, t, generator=gen).to(dev)
        cbs = [(torch.randn(k, dg, generator=gen) * 0.75 ** (i // g_)).to(dev) for i in range(s * g_)]
        hn = ops.codebook_half_norms(cbs)
        pack = ops.tc_pack_codebooks(cbs)
        fl = ops.ACQ_STE if g_ > 1 else 0
        ref, _, _, _ = ops.rvq_search(x, cbs, s, g_, half_norms=hn, flags=fl, impl=_lib.ACQ_IMPL_SIMT)
        first = None
        for _ in range(3):
            tc, _, _, _ = ops.rvq_search(x, cbs, s, g_, flags=fl, impl=_lib.ACQ_IMPL_TC, tc_pack=pack)
            if first is None:
                first = tc.clone()
            assert torch.equal(tc, first), f"non-deterministic: {(b, t, dg, g_, k, s)}"
        bad = int((tc != ref).any(dim=0).sum())
        tot += b * t
        diff += bad
        assert bad <= max(2, b * t // 500), (b, t, dg, g_, k, s, bad)
    print(f"[stress {variant} {'small' if small else 'multi-tile'}] {tot} frames, {diff} differ from SIMT "
          f"(near-ties + downstream)")


@pytest.mark.parametrize("cluster", [1, 2])
def test_single_stage_small_ring_many_undecided(dev, configure, cluster):
    """Single-stage calls settle their undecided frames through a per-CTA record ring; at D = 64 the ring holds
    exactly one tile of records.  Near-duplicate codewords make most frames undecided (the filter cannot separate
    them), so every tile fills the ring: the codes must still equal the SIMT kernel's up to float64 near-ties, launch
    after launch (round 2 shipped a back-pressure test against the reserved instead of the published record
    count, which deadlocked here once in a few runs)."""
    from academicodec_b200 import _lib, ops
    configure(1, cluster, 0)
    gen = torch.Generator(device="cpu").manual_seed(3)
    # (the last three take the wide layout -- streamed x, D % 128 == 0 -- with 1, 2 and 4 half slots per loader warp)
    for (b, t, d, k) in [(40, 127, 64, 256), (9, 750, 64, 1024), (300, 50, 128, 512),
                         (6, 1000, 128, 1024), (5, 2048, 256, 512), (3, 1500, 512, 256)]:
        base = torch.randn(k // 4, d, generator=gen)
        cb = (base.repeat_interleave(4, 0) * (1.0 + 2e-4 * torch.randn(k, 1, generator=gen))).to(dev)   # clusters of 4 near-duplicates
        x = torch.randn(b, d, t, generator=gen).to(dev)
        hn = ops.codebook_half_norms([cb])
        pack = ops.tc_pack_codebooks([cb])
        ref, _, _, _ = ops.rvq_search(x, [cb], 1, 1, half_norms=hn, impl=_lib.ACQ_IMPL_SIMT)
        first = None
        for _ in range(6):
            tc, _, _, _ = ops.rvq_search(x, [cb], 1, 1, impl=_lib.ACQ_IMPL_TC, tc_pack=pack)
            torch.cuda.synchronize()
            if first is None:
                first = tc.clone()
            assert torch.equal(tc, first)
        # exact float64 scores decide what a disagreement is
        bad = (tc != ref).view(-1).nonzero().view(-1)
        if bad.numel():
            xr = x.transpose(1, 2).reshape(-1, d)[bad].double()
            def sc(codes):
                e = cb.double()[codes.view(-1)[bad]]
                return (xr * e).sum(1) - 0.5 * (e * e).sum(1)
            gap = (sc(ref) - sc(tc)).abs() / (xr.norm(dim=1) * cb.double().norm(dim=1).max())
            assert float(gap.max()) < 1e-6, (b, t, d, k, int(bad.numel()), float(gap.max()))
        print(f"[small ring cl{cluster}] {(b, t, d, k)}: {int(bad.numel())} of {b * t} frames differ from SIMT (float64 near-ties)")


def test_auto_choice_guards_heterogeneous_codebooks(dev, configure):
    """Automatic kernel choice (variant 0): a table whose codewords differ widely in norm -- what EMA training
    produces: a few dead codes at their initial norm, the live ones contracted towards cluster means -- makes the
    single-product filter's bound useless (every row falls back to exact scores of all K codewords: 85 ms
    instead of 2 ms measured).  The pack flags such tables on the device and the three-product kernel, launched
    right behind, takes the call.  Codes must be right either way, and the call must not be slow."""
    from academicodec_b200 import _lib, ops
    gen = torch.Generator(device="cpu").manual_seed(3)
    b, d, t, k, s = 96, 512, 200, 1024, 3
    x = torch.randn(b, d, t, generator=gen).to(dev)

    def tables(hetero):
        out = []
        for i in range(s):
            w = torch.randn(k, d, generator=gen) * 0.8 ** i
            if hetero:
                w[64:] *= 0.04                     # 94 % of the codewords 25x smaller than the rest
            out.append(w.to(dev))
        return out

    def ms(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for hetero in (False, True):
        cbs = tables(hetero)
        pack = ops.tc_pack_codebooks(cbs)
        hn = ops.codebook_half_norms(cbs)
        ref, _, _, _ = ops.rvq_search(x, cbs, s, half_norms=hn, impl=_lib.ACQ_IMPL_SIMT)
        configure(0, 0, 0)
        run = lambda: ops.rvq_search(x, cbs, s, impl=_lib.ACQ_IMPL_TC, tc_pack=pack)[0]        # noqa: E731
        auto = run()
        t_auto = ms(run)
        configure(3, 1, 0)
        t3 = ms(run)
        bad = int((auto != ref).any(dim=0).sum())
        print(f"[auto choice hetero={hetero}] automatic {t_auto:.3f} ms, three-product {t3:.3f} ms, "
              f"{bad} of {b * t} frames differ from SIMT")
        assert bad <= max(2, b * t // 500)
        assert t_auto < 2.0 * t3 + 0.05, (t_auto, t3)
