"""Oracle: residual vector quantization (Encodec / SoundStream).  TEST INFRASTRUCTURE ONLY.

Functional restatement of the reference modules
  academicodec/quantization/core_vq.py  (EuclideanCodebook, VectorQuantization,
                                         ResidualVectorQuantization)
  academicodec/quantization/vq.py       (ResidualVectorQuantizer)
State is carried in plain dicts of tensors instead of nn.Module buffers.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]  # keys: embed [K,D], embed_avg [K,D], cluster_size [K], inited [1]


# ------------------------------------------------------------------ single Euclidean codebook
def nearest_codeword(flat: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """core_vq.py:175-180 -- negative squared distance, argmax, first index on ties.
    flat [N,D] fp32 contiguous, embed [K,D] -> [N] int64."""
    et = embed.t()
    neg_d = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ et + et.pow(2).sum(0, keepdim=True))
    return neg_d.max(dim=-1).indices


def lookup(indices: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """core_vq.py:185-187."""
    return F.embedding(indices, embed)


def _to_frames_last(x_bdt: torch.Tensor) -> torch.Tensor:
    """core_vq.py:285 ('b d n -> b n d') then :171-173 ('... d -> (...) d', a contiguous copy)."""
    return x_bdt.transpose(1, 2)


def layer_encode(x_bdt: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """VectorQuantization.encode (core_vq.py:284-288) -> EuclideanCodebook.encode (:189-197)."""
    x = _to_frames_last(x_bdt)
    flat = x.reshape(-1, x.shape[-1])
    return nearest_codeword(flat, embed).view(*x.shape[:-1])


def layer_decode(indices_bt: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """VectorQuantization.decode (core_vq.py:290-294): gather then 'b n d -> b d n'."""
    return lookup(indices_bt, embed).transpose(1, 2)


# ------------------------------------------------------------------ residual stack
def rvq_encode(x_bdt: torch.Tensor, embeds: List[torch.Tensor], n_q: Optional[int] = None,
               st: Optional[int] = None) -> torch.Tensor:
    """ResidualVectorQuantization.encode (core_vq.py:348-362).  Note `st` starts from the
    raw input: stages below `st` are *not* subtracted."""
    residual = x_bdt
    n_q = n_q or len(embeds)
    st = st or 0
    out = []
    for embed in embeds[st:n_q]:
        idx = layer_encode(residual, embed)
        residual = residual - layer_decode(idx, embed)
        out.append(idx)
    return torch.stack(out)


def rvq_decode(codes_sbt: torch.Tensor, embeds: List[torch.Tensor]) -> torch.Tensor:
    """ResidualVectorQuantization.decode (core_vq.py:364-370): left-to-right fp32 sum from 0.0."""
    acc = torch.tensor(0.0)
    for i, idx in enumerate(codes_sbt):
        acc = acc + layer_decode(idx, embeds[i])
    return acc


def _ema(moving: torch.Tensor, new: torch.Tensor, decay: float) -> None:
    """core_vq.py:47-48."""
    moving.mul_(decay).add_(new, alpha=(1 - decay))


def codebook_forward(x_bnd: torch.Tensor, state: State, training: bool, decay: float = 0.99,
                     epsilon: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """EuclideanCodebook.forward (core_vq.py:203-227) for an already-initialised codebook.

    Dead-code expiry (:159-169, :217) is omitted on purpose: it rewrites `embed`, which is
    overwritten from `embed_avg` a few lines later (:224-225), so it leaves no trace in the
    state (SURVEY.md section 7 'dead-code expiry no-op', probed)."""
    shape, dtype = x_bnd.shape, x_bnd.dtype
    flat = x_bnd.reshape(-1, shape[-1])
    embed = state["embed"]
    k = embed.shape[0]
    ind = nearest_codeword(flat, embed)
    onehot = F.one_hot(ind, k).type(dtype)
    ind_bt = ind.view(*shape[:-1])
    quantize = lookup(ind_bt, embed)            # uses the PRE-update codebook (:212)
    if training:
        _ema(state["cluster_size"], onehot.sum(0), decay)
        embed_sum = flat.t() @ onehot
        _ema(state["embed_avg"], embed_sum.t(), decay)
        cs = state["cluster_size"]
        smoothed = (cs + epsilon) / (cs.sum() + k * epsilon) * cs.sum()
        state["embed"].copy_(state["embed_avg"] / smoothed.unsqueeze(1))
    return quantize, ind_bt


def layer_forward(x_bdt: torch.Tensor, state: State, training: bool, decay: float = 0.99,
                  epsilon: float = 1e-5, commitment_weight: float = 1.0):
    """VectorQuantization.forward (core_vq.py:296-315).  project_in/out are Identity in every
    recipe."""
    x = _to_frames_last(x_bdt)
    quantize, ind = codebook_forward(x, state, training, decay, epsilon)
    if training:
        quantize = x + (quantize - x).detach()
    loss = torch.tensor([0.0], requires_grad=training)
    if training and commitment_weight > 0:
        loss = loss + F.mse_loss(quantize.detach(), x) * commitment_weight
    return quantize.transpose(1, 2), ind, loss


def rvq_forward(x_bdt: torch.Tensor, states: List[State], n_q: Optional[int] = None,
                training: bool = False, decay: float = 0.99, epsilon: float = 1e-5):
    """ResidualVectorQuantization.forward (core_vq.py:328-346)
    -> (quantized_out [B,D,T], indices [S,B,T], losses [S,1])."""
    quantized_out = 0.0
    residual = x_bdt
    losses, indices = [], []
    n_q = n_q or len(states)
    for st in states[:n_q]:
        q, idx, loss = layer_forward(residual, st, training, decay, epsilon)
        residual = residual - q
        quantized_out = quantized_out + q
        indices.append(idx)
        losses.append(loss)
    return quantized_out, torch.stack(indices), torch.stack(losses)


# ------------------------------------------------------------------ user-facing wrapper (vq.py)
def bandwidth_per_quantizer(bins: int, frame_rate: int) -> float:
    """vq.py:98-101 (the argument the reference calls sample_rate is the frame rate)."""
    return math.log2(bins) * frame_rate / 1000


def num_quantizers_for_bandwidth(n_q: int, bins: int, frame_rate: int,
                                 bandwidth: Optional[float] = None) -> int:
    """vq.py:88-96."""
    per_q = bandwidth_per_quantizer(bins, frame_rate)
    if bandwidth and bandwidth > 0.0:
        n_q = int(max(1, math.floor(bandwidth / per_q)))
    return n_q


def quantizer_forward(x_bdt, states, bins, frame_rate, bandwidth=None, training=False,
                      decay=0.99):
    """ResidualVectorQuantizer.forward (vq.py:67-86) -> (quantized, codes, bw, penalty)."""
    per_q = bandwidth_per_quantizer(bins, frame_rate)
    n_q = num_quantizers_for_bandwidth(len(states), bins, frame_rate, bandwidth)
    quantized, codes, commit = rvq_forward(x_bdt, states, n_q, training, decay)
    bw = torch.tensor(n_q * per_q).to(x_bdt)
    return quantized, codes, bw, torch.mean(commit)


# ------------------------------------------------------------------ k-means init (next row 8f-1)
def kmeans_from_means(samples: torch.Tensor, means: torch.Tensor, num_iters: int):
    """Lloyd iterations of core_vq.py:77-93 starting from given `means` (the reference draws
    them with torch.randperm, :61-69; the draw is passed in so both sides share it)."""
    k, dim = means.shape
    bins = None
    for _ in range(num_iters):
        diffs = samples.unsqueeze(1) - means.unsqueeze(0)
        dists = -(diffs ** 2).sum(dim=-1)
        buckets = dists.max(dim=-1).indices
        bins = torch.bincount(buckets, minlength=k)
        zero = bins == 0
        clamped = bins.masked_fill(zero, 1)
        new_means = buckets.new_zeros(k, dim, dtype=samples.dtype)
        new_means.scatter_add_(0, buckets.unsqueeze(1).expand(-1, dim), samples)
        new_means = new_means / clamped[..., None]
        means = torch.where(zero[..., None], means, new_means)
    return means, bins


def make_states(embeds) -> List[State]:
    """Fresh EMA state for explicitly initialised codebooks: embed_avg = embed.clone()
    (core_vq.py:134-137), cluster_size zeros, inited True."""
    out = []
    for e in embeds:
        e = torch.as_tensor(e).clone()
        out.append(dict(embed=e, embed_avg=e.clone(), cluster_size=torch.zeros(e.shape[0]),
                        inited=torch.ones(1)))
    return out
