"""Generate the golden fixtures by running the UNMODIFIED live reference on CPU.

    python tests/golden/make_golden.py            (dev container only; needs /root/reference)

The reference (jacquelm/AcademiCodec) is imported read-only from $ACADEMICODEC_REF or
/root/reference.  Inputs are rebuilt from seeds (tests/cases.py + academicodec_b200/synth.py)
so the .npz files hold only what the reference *returned*: codes, quantized latents, losses
and updated EMA buffers.  The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("ACADEMICODEC_REF", "/root/reference")
sys.path.insert(0, REF)

# hificodec/models.py imports academicodec.utils which imports matplotlib (absent here)
for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        sys.modules[name] = m

from academicodec.quantization import ResidualVectorQuantizer  # noqa: E402  (reference)
from academicodec.quantization import core_vq as ref_core_vq  # noqa: E402  (reference)
from academicodec.models.hificodec.models import Quantizer as RefQuantizer  # noqa: E402

from tests import cases  # noqa: E402

ROW_STRIDE = 64   # buffers are sampled every 64th codeword row (plus whole-buffer checksums)


def _load_rvq(case, cb):
    q = ResidualVectorQuantizer(dimension=case["D"], n_q=case["n_q"], bins=case["bins"],
                                kmeans_init=False)
    for i, layer in enumerate(q.vq.layers):
        c = layer._codebook
        c.embed.data.copy_(cb[i])
        c.embed_avg.data.copy_(cb[i])
        c.cluster_size.data.zero_()
        c.inited.data.fill_(1.0)
    return q


def gen_rvq(name, case, x, cb, out):
    fr = case["frame_rate"]
    q = _load_rvq(case, cb).eval()
    with torch.no_grad():
        codes = q.encode(x, fr)
        out[f"{name}/codes"] = codes.numpy().astype(np.int16)
        out[f"{name}/decode"] = q.decode(codes).numpy()
        # bandwidth-limited: pick a bandwidth that yields about half the stages
        per_q = q.get_bandwidth_per_quantizer(fr)
        bw = per_q * max(1, case["n_q"] // 2) + 1e-3
        out[f"{name}/bw"] = np.float64(bw)
        codes_bw = q.encode(x, fr, bw)
        out[f"{name}/codes_bw"] = codes_bw.numpy().astype(np.int16)
        if case["n_q"] >= 3:
            out[f"{name}/codes_st2"] = q.encode(x, fr, None, 2).numpy().astype(np.int16)
        qz, cz, bwt, pen = q(x, fr, bw)
        out[f"{name}/fwd_eval_quantized"] = qz.numpy()
        out[f"{name}/fwd_eval_codes"] = cz.numpy().astype(np.int16)
        out[f"{name}/fwd_eval_bw"] = bwt.numpy()
        out[f"{name}/fwd_eval_penalty"] = pen.numpy()
    # train mode: two EMA steps with the reference's default dead-code threshold (2)
    q = _load_rvq(case, cb).train()
    torch.manual_seed(0)
    for step in range(2):
        xs = x if step == 0 else x.flip(0) * 0.5
        qz, cz, bwt, pen = q(xs, fr)
        out[f"{name}/train{step}_quantized"] = qz.detach().numpy()
        out[f"{name}/train{step}_codes"] = cz.numpy().astype(np.int16)
        out[f"{name}/train{step}_penalty"] = pen.detach().numpy()
    # keep the fixtures small: a strided row sample + float64 checksums of every buffer
    for i, layer in enumerate(q.vq.layers):
        c = layer._codebook
        out[f"{name}/train_cluster_size{i}"] = c.cluster_size.numpy().copy()
        for key, buf in (("embed", c.embed), ("embed_avg", c.embed_avg)):
            a = buf.numpy()
            out[f"{name}/train_{key}{i}_rows"] = a[::ROW_STRIDE].copy()
            out[f"{name}/train_{key}{i}_sums"] = np.array(
                [a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])


def gen_ties(out):
    x, cb = cases.tie_inputs()
    case = dict(D=cb.shape[2], n_q=cb.shape[0], bins=cb.shape[1])
    q = _load_rvq(case, cb).eval()
    with torch.no_grad():
        codes = q.encode(x, 100)
        out["ties/codes"] = codes.numpy().astype(np.int16)
        out["ties/decode"] = q.decode(codes).numpy()


def gen_grvq(name, case, x, w, out):
    h = types.SimpleNamespace(n_code_groups=case["G"], n_codes=case["n_codes"],
                              codebook_loss_lambda=1.0, commitment_loss_lambda=0.25)
    q = RefQuantizer(h)
    with torch.no_grad():
        for g in range(case["G"]):
            q.quantizer_modules[g].embedding.weight.copy_(w[0][g])
            q.quantizer_modules2[g].embedding.weight.copy_(w[1][g])
        qo, loss, ids = q(x)
        out[f"{name}/quantized"] = qo.numpy()
        out[f"{name}/loss"] = loss.numpy()
        codes = torch.stack(ids, -1).reshape(x.shape[0], x.shape[2], -1)   # vqvae.py:41-45
        out[f"{name}/codes"] = codes.numpy().astype(np.int16)
        out[f"{name}/embed"] = q.embed(codes).numpy()
    # gradients (autograd contract)
    xg = x.clone().requires_grad_(True)
    qo, loss, _ = q(xg)
    (qo.square().mean() + 10.0 * loss).backward()
    out[f"{name}/grad_x"] = xg.grad.numpy()
    for key, mod in (("grad_w00", q.quantizer_modules[0]), ("grad_w10", q.quantizer_modules2[0])):
        a = mod.embedding.weight.grad.numpy()
        out[f"{name}/{key}_rows"] = a[::ROW_STRIDE].copy()
        out[f"{name}/{key}_sums"] = np.array([a.astype(np.float64).sum(),
                                              np.abs(a.astype(np.float64)).sum()])


def gen_kmeans(out):
    samples, k, iters = cases.kmeans_inputs()
    torch.manual_seed(77)
    pick = torch.randperm(samples.shape[0])[:k]
    torch.manual_seed(77)
    means, bins = ref_core_vq.kmeans(samples, k, iters)
    out["kmeans/pick"] = pick.numpy()
    out["kmeans/means"] = means.numpy()
    out["kmeans/bins"] = bins.numpy()


def gen_rvq_grad(out):
    """RVQ train-mode autograd contract: straight-through + stage-0 commit-loss gradient."""
    case = cases.RVQ_CASES["odd_dims"]
    x, cb = cases.rvq_inputs(case)
    q = _load_rvq(case, cb).train()
    xg = x.clone().requires_grad_(True)
    qz, _, _, pen = q(xg, case["frame_rate"])
    (qz.square().mean() + 3.0 * pen).backward()
    out["rvq_grad/grad_x"] = xg.grad.numpy()


def gen_bitpack(out):
    """Code wire format: the reference's own BitPacker / BitUnpacker (binary.py:54-123)."""
    import io
    from academicodec import binary as ref_binary
    for bits, n in cases.BITPACK_CASES:
        vals = cases.bitpack_values(bits, n)
        fo = io.BytesIO()
        packer = ref_binary.BitPacker(bits, fo)
        for v in vals:
            packer.push(int(v))
        packer.flush()
        data = fo.getvalue()
        out[f"bitpack/{bits}_{n}"] = np.frombuffer(data, dtype=np.uint8).copy()
        unp = ref_binary.BitUnpacker(bits, io.BytesIO(data))
        back = np.array([unp.pull() for _ in range(n)], dtype=np.int64)
        assert np.array_equal(back, vals)


def main():
    torch.set_num_threads(1)          # deterministic reduction order for the fixtures
    out = {}
    for name, case in cases.RVQ_CASES.items():
        x, cb = cases.rvq_inputs(case)
        gen_rvq(name, case, x, cb, out)
    gen_ties(out)
    for name, case in cases.GRVQ_CASES.items():
        x, w = cases.grvq_inputs(case)
        gen_grvq(name, case, x, w, out)
    gen_kmeans(out)
    gen_rvq_grad(out)
    gen_bitpack(out)
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
