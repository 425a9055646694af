"""Oracle: group-residual VQ (HiFi-Codec).  TEST INFRASTRUCTURE ONLY.

Functional restatement of `Quantizer_module` / `Quantizer` in the reference
academicodec/models/hificodec/models.py:430-535.  Codebooks are passed as
`weights[stage][group]`, each `[n_codes, 512 // n_groups]` (the reference hard-codes the
512-channel latent, models.py:448,465).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F

CHANNELS = 512  # models.py:448,450,465-466


def group_nearest(x2d: torch.Tensor, w: torch.Tensor):
    """Quantizer_module.forward (models.py:436-442): squared distance, argmin (first on ties),
    embedding lookup."""
    d = torch.sum(x2d ** 2, 1, keepdim=True) + torch.sum(w ** 2, 1) - 2 * torch.matmul(x2d, w.T)
    idx = torch.argmin(d, 1)
    return F.embedding(idx, w), idx


def one_stage(xin_bct: torch.Tensor, group_weights: Sequence[torch.Tensor],
              codebook_loss_lambda: float, commitment_loss_lambda: float):
    """Quantizer.for_one_step (models.py:463-492); both branches of the reference are the same
    code over a different ModuleList, so one function serves both stages."""
    n_groups = len(group_weights)
    xin = xin_bct.transpose(1, 2)
    x = xin.reshape(-1, CHANNELS)
    parts = torch.split(x, CHANNELS // n_groups, dim=-1)
    zs, ids = [], []
    for part, w in zip(parts, group_weights):
        z, i = group_nearest(part, w)
        zs.append(z)
        ids.append(i)
    z_q = torch.cat(zs, -1).reshape(xin.shape)
    loss = codebook_loss_lambda * torch.mean((z_q - xin.detach()) ** 2) \
        + commitment_loss_lambda * torch.mean((z_q.detach() - xin) ** 2)
    z_q = xin + (z_q - xin).detach()          # straight-through, applied in eval too
    return z_q.transpose(1, 2), loss, ids


def grvq_forward(xin_bct: torch.Tensor, weights: List[List[torch.Tensor]],
                 codebook_loss_lambda: float = 1.0, commitment_loss_lambda: float = 0.25):
    """Quantizer.forward (models.py:494-508)
    -> (quantized_out [B,512,T], loss 0-d, [idx_s0g0, idx_s0g1, ..., idx_s1g0, ...] each [B*T])."""
    quantized_out = 0.0
    residual = xin_bct
    losses, indices = [], []
    for stage_weights in weights:
        q, loss, ids = one_stage(residual, stage_weights, codebook_loss_lambda,
                                 commitment_loss_lambda)
        residual = residual - q
        quantized_out = quantized_out + q
        indices.extend(ids)
        losses.append(loss)
    return quantized_out, torch.mean(torch.stack(losses)), indices


def grvq_embed(codes_btc: torch.Tensor, weights: List[List[torch.Tensor]]) -> torch.Tensor:
    """Quantizer.embed (models.py:510-535): codes [B,T,2G] in the order s0g0,s0g1,..,s1g0,..
    -> [B,512,T]; per stage concatenate the group lookups, sum the stages from 0.0."""
    n_groups = len(weights[0])
    acc = torch.tensor(0.0)
    cols = torch.split(codes_btc, 1, 2)
    for s, stage_weights in enumerate(weights):
        parts = [F.embedding(cols[s * n_groups + g].squeeze(-1), stage_weights[g])
                 for g in range(n_groups)]
        acc = acc + torch.cat(parts, -1)
    return acc.transpose(1, 2)
