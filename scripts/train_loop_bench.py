"""main_launch.py's training loop body with the quantizer swapped (VERDICT r01 missing #1).

    python scripts/train_loop_bench.py [--quantizer ref|ours|both] [--steps 10] [--batch 8] [--seconds 1.0]
    python -m torch.distributed.run --nproc-per-node N ... scripts/train_loop_bench.py ...     (DDP, NCCL)

Runs the loop of /root/reference/academicodec/models/encodec/main_launch.py:285-327 -- two forwards of the
generator per batch (:289-291), generator update against the STFT / multi-period / multi-scale discriminators
with the reference's own `loss_g`, then the discriminator update with `loss_dis` -- built from the UNMODIFIED
reference classes in baseline/_ref (SoundStream n_filters=32, D=512, ratios 6 5 4 2 = Encodec_24k_240d, the three
discriminators, losses, AdamW(3e-4, betas (0.5, 0.9)) as in :239-247) on synthetic audio.

  --quantizer ref    the stock model; under DDP with the default broadcast_buffers=True (:199-204), i.e. rank 0's
                     codebook buffers are re-broadcast before every forward (50 MB per forward at n_q=12, D=512)
  --quantizer ours   `swap_quantizer(soundstream)` before the DDP wrap; DDP(broadcast_buffers=False): the EMA
                     statistics are all-reduced inside the quantizer and every rank applies the same update

Prints per arm: ms per step (CUDA events, max over ranks), the commitment-loss trajectory, and under DDP whether
the codebooks of all ranks are bit-identical after the last step.  `run()` is imported by tests/test_gpu_models.py.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _reference():
    from baseline import install_ref
    pkg = install_ref.load()
    if pkg is None:
        raise RuntimeError("baseline/_ref is not installed (python baseline/install_ref.py in the dev container)")
    return pkg


def _args(device, sr=24000):
    # main_launch.py:60-100 defaults
    return types.SimpleNamespace(LAMBDA_ADV=1.0, LAMBDA_FEAT=1.0, LAMBDA_REC=1.0, LAMBDA_COM=1000.0, LAMBDA_WAV=100.0,
                                 discriminator_iter_start=500, sr=sr, device=device)


def _init_codebooks(q, seed=7, shrink=0.8):
    g = torch.Generator().manual_seed(seed)
    for i, layer in enumerate(q.vq.layers):
        cb = layer._codebook
        w = torch.randn(cb.embed.shape, generator=g) * (shrink ** i)
        cb.embed.data.copy_(w)
        cb.embed_avg.data.copy_(w)
        cb.cluster_size.data.fill_(1.0)
        cb.inited.data.fill_(1.0)


def run(quantizer="ours", steps=10, warmup=3, batch=8, seconds=1.0, device=None, seed=1, verbose=True,
        start_step=0):
    _reference()
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from academicodec.models.encodec.net3 import SoundStream
    from academicodec.models.encodec.msstftd import MultiScaleSTFTDiscriminator
    from academicodec.models.soundstream.models import MultiPeriodDiscriminator, MultiScaleDiscriminator
    from academicodec.models.encodec.loss import loss_g, loss_dis
    from academicodec_b200.codec import swap_quantizer

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    device = device or torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    use_cuda = device.type == "cuda"
    args = _args(device)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    soundstream = SoundStream(n_filters=32, D=512, ratios=[6, 5, 4, 2], sample_rate=24000,
                              target_bandwidths=[1, 1.5, 2, 4, 6, 12])
    _init_codebooks(soundstream.quantizer)
    msd = MultiScaleDiscriminator()
    mpd = MultiPeriodDiscriminator()
    stft_disc = MultiScaleSTFTDiscriminator(filters=32)
    soundstream.to(device); msd.to(device); mpd.to(device); stft_disc.to(device)
    if quantizer == "ours":
        swap_quantizer(soundstream)
    if distributed:
        # main_launch.py:199-204 wraps with the defaults (broadcast_buffers=True); with the all-reducing
        # quantizer the buffers never diverge, so the broadcast is switched off
        bb = quantizer != "ours"
        soundstream = DDP(soundstream, device_ids=[device.index], broadcast_buffers=bb)
        msd = DDP(msd, device_ids=[device.index])
        mpd = DDP(mpd, device_ids=[device.index])
        stft_disc = DDP(stft_disc, device_ids=[device.index])
    core = soundstream.module if distributed else soundstream
    optimizer_g = torch.optim.AdamW(soundstream.parameters(), lr=3e-4, betas=(0.5, 0.9))
    import itertools
    optimizer_d = torch.optim.AdamW(itertools.chain(stft_disc.parameters(), msd.parameters(), mpd.parameters()),
                                    lr=3e-4, betas=(0.5, 0.9))
    soundstream.train(); stft_disc.train(); msd.train(); mpd.train()
    n = int(24000 * seconds)
    gen = torch.Generator().manual_seed(100 + rank)
    out = {"commit": [], "loss_g": [], "loss_d": []}
    ev0 = ev1 = None
    global_step = start_step
    for it in range(warmup + steps):
        if it == warmup:
            if use_cuda:
                torch.cuda.synchronize(device)
            if distributed:
                dist.barrier()
            if use_cuda:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            t_host = time.perf_counter()
        t = torch.arange(n) / 24000.0
        x = (0.4 * torch.sin(2 * np.pi * 220.0 * t)[None] * torch.rand(batch, 1, generator=gen)
             + 0.1 * torch.randn(batch, n, generator=gen)).unsqueeze(1).to(device)
        global_step += 1
        for optimizer_idx in [0, 1]:                                     # main_launch.py:288
            x_wav = x
            G_x, commit_loss, last_layer = soundstream(x_wav)            # :290 (twice per batch)
            if optimizer_idx == 0:
                y_disc_r, fmap_r = stft_disc(x_wav.contiguous())
                y_disc_gen, fmap_gen = stft_disc(G_x.contiguous())
                y_df_hat_r, y_df_hat_g, fmap_f_r, fmap_f_g = mpd(x_wav.contiguous(), G_x.contiguous())
                y_ds_hat_r, y_ds_hat_g, fmap_s_r, fmap_s_g = msd(x_wav.contiguous(), G_x.contiguous())
                total_loss_g, rec_loss, adv_g_loss, feat_loss, d_weight = loss_g(
                    commit_loss, x_wav, G_x, fmap_r, fmap_gen, y_disc_r, y_disc_gen, global_step, y_df_hat_r,
                    y_df_hat_g, y_ds_hat_r, y_ds_hat_g, fmap_f_r, fmap_f_g, fmap_s_r, fmap_s_g,
                    last_layer=last_layer, is_training=True, args=args)
                if it >= warmup:
                    out["commit"].append(float(commit_loss.item()))
                    out["loss_g"].append(float(total_loss_g.item()))
                optimizer_g.zero_grad()
                total_loss_g.backward()
                optimizer_g.step()
            else:
                y_disc_r_det, fmap_r_det = stft_disc(x.detach())
                y_disc_gen_det, fmap_gen_det = stft_disc(G_x.detach())
                y_df_hat_r, y_df_hat_g, fmap_f_r, fmap_f_g = mpd(x.detach(), G_x.detach())
                y_ds_hat_r, y_ds_hat_g, fmap_s_r, fmap_s_g = msd(x.detach(), G_x.detach())
                loss_d = loss_dis(y_disc_r_det, y_disc_gen_det, fmap_r_det, fmap_gen_det, y_df_hat_r, y_df_hat_g,
                                  fmap_f_r, fmap_f_g, y_ds_hat_r, y_ds_hat_g, fmap_s_r, fmap_s_g, global_step, args)
                if it >= warmup:
                    out["loss_d"].append(float(loss_d.item()) if torch.is_tensor(loss_d) else float(loss_d))
                if torch.is_tensor(loss_d) and loss_d.requires_grad:
                    optimizer_d.zero_grad()
                    loss_d.backward()
                    optimizer_d.step()
    if use_cuda:
        ev1.record()
        torch.cuda.synchronize(device)
        ms = torch.tensor([ev0.elapsed_time(ev1) / max(1, steps)], device=device)
    else:
        ms = torch.tensor([(time.perf_counter() - t_host) * 1e3 / max(1, steps)])
    same = None
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # are the codebooks of all ranks bit-identical after training?
        flat = torch.cat([l._codebook.embed.reshape(-1) for l in core.quantizer.vq.layers])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    out.update(quantizer=quantizer, ms_per_step=float(ms.item()), world=world, batch_per_gpu=batch, seconds=seconds,
               codebooks_identical_across_ranks=same,
               broadcast_buffers=(quantizer != "ours") if distributed else None)
    if verbose and rank == 0:
        print(json.dumps(out), flush=True)
    del soundstream, msd, mpd, stft_disc, optimizer_g, optimizer_d
    if use_cuda:
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quantizer", default="both", choices=["ref", "ours", "both"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=1.0)
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for q in (["ref", "ours"] if a.quantizer == "both" else [a.quantizer]):
        run(quantizer=q, steps=a.steps, warmup=a.warmup, batch=a.batch, seconds=a.seconds, device=dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
