"""The reference's own model classes with the quantizer swapped (BASELINE cfg4).

`baseline/_ref` holds the UNMODIFIED reference package (baseline/install_ref.py copies it from /root/reference
in the dev container; it is git-ignored and travels to the GPU box with the snapshot).  These tests build the
reference's models with their stock constructors, deep-copy them, replace ONLY the quantizer of the copy with
`academicodec_b200.codec.swap_quantizer`, and compare the two models on the same GPU:

  * SoundStream (net3.py:12-61; SEANet encoder -> RVQ n_q=12 -> SEANet decoder, Encodec_24k_240d = cfg4):
    encode() codes index for index (fp64-adjudicated where the reference's own fp32 rounding decides a near-tie),
    quantizer.decode() bit-exact, decode() waveform identical, eval forward; a training forward's commitment
    loss and EMA buffers.
  * HiFi-Codec Encoder / Quantizer / Generator behind the reference's VQVAE.encode / forward (vqvae.py:31-45).
  * Ten steps of main_launch.py's training loop body (:285-327; generator + the three discriminators, the
    reference's own losses) with the module swapped: finite, decreasing-or-equal bookkeeping, and the first
    step's commitment loss equal to the reference's.

TF32 is off and cuDNN is deterministic, so both models see bit-identical encoder outputs.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ref():
    from baseline import install_ref
    pkg = install_ref.load()
    if pkg is None:
        pytest.skip("baseline/_ref (the unmodified reference package) is not installed")
    return pkg


def _init_rvq_codebooks(q, seed=7, shrink=0.8):
    """Give a reference ResidualVectorQuantizer trained-looking codebooks without its 50-iteration k-means
    (kmeans_init=True would run it inside the first training forward): stage s ~ N(0, shrink^2s)."""
    g = torch.Generator().manual_seed(seed)
    for i, layer in enumerate(q.vq.layers):
        cb = layer._codebook
        w = torch.randn(cb.embed.shape, generator=g) * (shrink ** i)
        cb.embed.data.copy_(w)
        cb.embed_avg.data.copy_(w)
        cb.cluster_size.data.fill_(1.0)
        cb.inited.data.fill_(1.0)


def _soundstream(ref, dev, seed=0):
    from academicodec.models.encodec.net3 import SoundStream
    torch.manual_seed(seed)
    # egs/Encodec_24k_240d/start.sh: --ratios 6 5 4 2 --target_bandwidths 1 1.5 2 4 6 12 --sr 24000
    m = SoundStream(n_filters=32, D=512, ratios=[6, 5, 4, 2], sample_rate=24000,
                    target_bandwidths=[1, 1.5, 2, 4, 6, 12])
    assert m.quantizer.n_q == 12 and m.frame_rate == 100
    _init_rvq_codebooks(m.quantizer)
    return m.to(dev)


def _twin(build, model):
    """A second instance with the same state (copy.deepcopy does not work on weight-normed modules)."""
    twin = build()
    twin.load_state_dict(model.state_dict())
    return twin.train(model.training)


def _audio(b, n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 24000.0
    x = 0.4 * torch.sin(2 * np.pi * 220.0 * t)[None] * torch.rand(b, 1, generator=g) + 0.1 * torch.randn(b, n, generator=g)
    return x.unsqueeze(1).to(dev)


def test_soundstream_encode_decode_with_swapped_quantizer(ref, dev):
    from academicodec_b200.codec import swap_quantizer
    from academicodec_b200.quantization import ResidualVectorQuantizer
    from oracle import adjudicate
    model_ref = _soundstream(ref, dev).eval()
    model_new = swap_quantizer(_twin(lambda: _soundstream(ref, dev), model_ref)).eval()
    assert isinstance(model_new.quantizer, ResidualVectorQuantizer)
    assert type(model_ref.quantizer).__module__.startswith("academicodec.quantization")
    x = _audio(4, 24000, 1, dev)                        # 4 clips x 1 s
    with torch.no_grad():
        e_ref = model_ref.encoder(x)
        e_new = model_new.encoder(x)
        assert torch.equal(e_ref, e_new)                # same conv nets, deterministic
        for bw in (None, 6, 1.5):
            codes_ref = model_ref.encode(x, target_bw=bw)
            codes_new = model_new.encode(x, target_bw=bw)
            assert codes_ref.shape == codes_new.shape and codes_new.dtype == torch.int64
            embeds = [l._codebook.embed.cpu() for l in model_ref.quantizer.vq.layers]
            rep = adjudicate.compare_rvq_codes(e_ref.cpu(), embeds, codes_ref.cpu().numpy(), codes_new.cpu().numpy())
            print(f"[soundstream bw={bw}] codes={rep['total']} identical={rep['identical']} near_tie={rep['near_tie']} "
                  f"downstream={rep['downstream']} hard={rep['hard_mismatch']}")
            assert rep["hard_mismatch"] == 0, rep["hard_examples"]
            assert rep["diverged_frames"] <= 2, rep
            # decode: the quantizer's gather-accumulate is bit-exact, so the waveform is identical
            q_ref = model_ref.quantizer.decode(codes_ref)
            q_new = model_new.quantizer.decode(codes_ref)
            assert torch.equal(q_ref, q_new)
            assert torch.equal(model_ref.decode(codes_ref), model_new.decode(codes_ref))
        # encode from a later stage (net3.py:46-57 `st`)
        c_ref = model_ref.encode(x, target_bw=6, st=2)
        c_new = model_new.encode(x, target_bw=6, st=2)
        assert c_ref.shape == c_new.shape
        assert (c_ref != c_new).any(0).float().mean().item() <= 0.01


def test_soundstream_forward_eval_and_train(ref, dev):
    import random
    from academicodec_b200.codec import swap_quantizer
    model_ref = _soundstream(ref, dev, seed=3)
    model_new = swap_quantizer(_twin(lambda: _soundstream(ref, dev, seed=3), model_ref))
    x = _audio(4, 24000, 2, dev)
    # eval forward: (waveform, commit_loss, None); the bandwidth is drawn with python's random (net3.py:41)
    model_ref.eval(); model_new.eval()
    with torch.no_grad():
        random.seed(11); o_ref, l_ref, n_ref = model_ref(x)
        random.seed(11); o_new, l_new, n_new = model_new(x)
    assert n_ref is None and n_new is None
    assert o_ref.shape == o_new.shape
    assert torch.allclose(o_ref, o_new, rtol=1e-4, atol=1e-5) or (o_ref - o_new).abs().mean() < 1e-4
    # training forward: commitment loss, gradient w.r.t. the encoder, EMA buffers after the step
    model_ref.train(); model_new.train()
    random.seed(5); o_ref, l_ref, _ = model_ref(x)
    random.seed(5); o_new, l_new, _ = model_new(x)
    assert torch.allclose(l_ref, l_new, rtol=1e-4, atol=1e-7), (float(l_ref), float(l_new))
    (o_ref.abs().mean() + l_ref).backward()
    (o_new.abs().mean() + l_new).backward()
    (name, p_ref), (_, p_new) = next(zip(model_ref.encoder.named_parameters(), model_new.encoder.named_parameters()))
    g_ref, g_new = p_ref.grad, p_new.grad
    assert g_ref is not None and g_new is not None, name
    assert torch.allclose(g_ref, g_new, rtol=2e-3, atol=1e-6)
    used = 0
    for lr_, ln_ in zip(model_ref.quantizer.vq.layers, model_new.quantizer.vq.layers):
        a, b = lr_._codebook, ln_._codebook
        if torch.equal(a.cluster_size, torch.ones_like(a.cluster_size)):
            continue                                      # stage not used at the drawn bandwidth
        used += 1
        # (a near-tie resolved differently moves one frame between two clusters: allow a handful of rows)
        def rows_off(u, v, rtol, atol):
            bad = ~torch.isclose(u, v, rtol=rtol, atol=atol)
            return int(bad.reshape(bad.shape[0], -1).any(1).sum())
        assert rows_off(a.cluster_size[:, None], b.cluster_size[:, None], 1e-4, 1e-5) <= 4
        assert rows_off(a.embed_avg, b.embed_avg, 1e-4, 1e-5) <= 4
        assert rows_off(a.embed, b.embed, 1e-3, 1e-5) <= 4
    assert used >= 1


def _hifi_h():
    # egs/HiFi-Codec-16k-320d/config_16k_320d.json (the keys the three modules read)
    return types.SimpleNamespace(
        resblock="1", upsample_rates=[8, 5, 4, 2], upsample_kernel_sizes=[16, 11, 8, 4],
        upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
        resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], n_code_groups=2, n_codes=1024,
        codebook_loss_lambda=1.0, commitment_loss_lambda=0.25, sampling_rate=16000)


def test_hificodec_vqvae_encode_forward(ref, dev):
    """The reference's VQVAE.encode / forward (vqvae.py:31-45) run unmodified on an object whose quantizer is
    ours: VQVAE.__init__ needs a checkpoint file, so the three modules are attached to a bare instance."""
    from academicodec.models.hificodec.models import Encoder, Generator, Quantizer as RefQuantizer
    from academicodec.models.hificodec.vqvae import VQVAE
    from academicodec_b200.codec import swap_quantizer
    from academicodec_b200.grvq import Quantizer
    h = _hifi_h()
    torch.manual_seed(2)

    def build():
        m = VQVAE.__new__(VQVAE)
        torch.nn.Module.__init__(m)
        m.h = h
        m.quantizer = RefQuantizer(h)
        m.generator = Generator(h)
        m.encoder = Encoder(h)
        return m.to(dev)
    vq_ref = build()
    with torch.no_grad():                              # trained-looking codebooks (init range is 1/1024)
        g = torch.Generator().manual_seed(9)
        for i, qm in enumerate(list(vq_ref.quantizer.quantizer_modules) + list(vq_ref.quantizer.quantizer_modules2)):
            w = torch.randn(qm.embedding.weight.shape, generator=g) * (0.3 if i < 2 else 0.15)
            qm.embedding.weight.copy_(w.to(dev))
    vq_ref = vq_ref.eval()
    vq_new = swap_quantizer(_twin(build, vq_ref)).eval()
    assert isinstance(vq_new.quantizer, Quantizer)
    wav = _audio(4, 16000, 4, dev).squeeze(1)           # [B, L]
    with torch.no_grad():
        codes_ref = vq_ref.encode(wav)                  # [B, T, 4]
        codes_new = vq_new.encode(wav)
        assert codes_ref.shape == codes_new.shape == (4, 50, 4)
        frac = (codes_ref != codes_new).any(-1).float().mean().item()
        print(f"[hificodec] frames with a differing code: {frac:.5f}")
        assert frac <= 0.01
        # stage-0 disagreements must be fp64 near-ties on the encoder output
        from oracle import adjudicate
        c = vq_ref.encoder(wav.unsqueeze(1))
        flat = c.transpose(1, 2).reshape(-1, 512).cpu()
        for gi in range(2):
            w = vq_ref.quantizer.quantizer_modules[gi].embedding.weight.detach().cpu()
            sub = flat[:, gi * 256:(gi + 1) * 256]
            a = codes_ref[..., gi].reshape(-1).cpu().numpy()
            b = codes_new[..., gi].reshape(-1).cpu().numpy()
            bad = np.nonzero(a != b)[0]
            if len(bad):
                ja = adjudicate.judge_choice(sub[bad], w, a[bad])
                jb = adjudicate.judge_choice(sub[bad], w, b[bad])
                assert (np.abs(ja["excess"] - jb["excess"]) <= ja["tol"]).all()
        # embed + generator: VQVAE.forward on the same codes
        assert torch.equal(vq_ref.quantizer.embed(codes_ref), vq_new.quantizer.embed(codes_ref))
        assert torch.equal(vq_ref(codes_ref), vq_new(codes_ref))


def test_training_loop_with_swapped_quantizer(ref, dev):
    """scripts/train_loop_bench.py runs main_launch.py's loop body (:285-327) with the reference's SoundStream,
    discriminators and losses; here: a few steps per arm on one GPU, commitment losses compared."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import train_loop_bench as tlb
    out_ref = tlb.run(quantizer="ref", steps=4, warmup=0, batch=2, seconds=0.5, device=dev, seed=1, verbose=False)
    out_new = tlb.run(quantizer="ours", steps=4, warmup=0, batch=2, seconds=0.5, device=dev, seed=1, verbose=False)
    for k in ("commit", "loss_g", "loss_d"):
        assert all(np.isfinite(v) for v in out_new[k]), out_new
    # step 0 sees identical weights and codebooks: same commitment loss (near-tie flips change it in the 5th digit)
    assert abs(out_ref["commit"][0] - out_new["commit"][0]) <= 1e-3 * abs(out_ref["commit"][0]) + 1e-7, (out_ref, out_new)
    assert abs(out_ref["loss_g"][0] - out_new["loss_g"][0]) <= 1e-3 * abs(out_ref["loss_g"][0]), (out_ref, out_new)
    # later steps follow separately updated (Adam, EMA) models: same trajectory within a loose band
    for a, b in zip(out_ref["commit"], out_new["commit"]):
        assert abs(a - b) <= 0.1 * abs(a) + 1e-6, (out_ref["commit"], out_new["commit"])
