"""User-facing residual vector quantizer: drop-in for the reference's
academicodec/quantization/vq.py `ResidualVectorQuantizer` (:27-121) -- same constructor
arguments, same `forward` 4-tuple, same `encode(x, sample_rate, bandwidth, st)` /
`decode(codes)` and the same state-dict keys (`vq.layers.{i}._codebook.*`)."""
from __future__ import annotations

import math
import typing as tp
from dataclasses import dataclass, field

import torch
from torch import nn

from .core_vq import ResidualVectorQuantization


@dataclass
class QuantizedResult:
    """Declared by the reference (vq.py:18-24) but not returned by it; kept for import parity."""
    quantized: torch.Tensor
    codes: torch.Tensor
    bandwidth: torch.Tensor
    penalty: tp.Optional[torch.Tensor] = None
    metrics: dict = field(default_factory=dict)


class ResidualVectorQuantizer(nn.Module):
    def __init__(self, dimension: int = 256, n_q: int = 8, bins: int = 1024, decay: float = 0.99,
                 kmeans_init: bool = True, kmeans_iters: int = 50,
                 threshold_ema_dead_code: int = 2):
        super().__init__()
        self.n_q = n_q
        self.dimension = dimension
        self.bins = bins
        self.decay = decay
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.vq = ResidualVectorQuantization(
            dim=dimension, codebook_size=bins, num_quantizers=n_q, decay=decay,
            kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
            threshold_ema_dead_code=threshold_ema_dead_code)

    # `sample_rate` is what the reference calls it; callers pass the FRAME rate (net3.py:42-43,55)
    def get_bandwidth_per_quantizer(self, sample_rate: int) -> float:
        return math.log2(self.bins) * sample_rate / 1000

    def get_num_quantizers_for_bandwidth(self, sample_rate: int,
                                         bandwidth: tp.Optional[float] = None) -> int:
        n_q = self.n_q
        if bandwidth and bandwidth > 0.0:
            n_q = int(max(1, math.floor(bandwidth / self.get_bandwidth_per_quantizer(sample_rate))))
        return n_q

    def forward(self, x: torch.Tensor, sample_rate: int, bandwidth: tp.Optional[float] = None):
        """x [B, D, T] -> (quantized [B, D, T], codes [n_q', B, T] int64, bandwidth 0-d, penalty 0-d)."""
        n_q = self.get_num_quantizers_for_bandwidth(sample_rate, bandwidth)
        quantized, codes, commit_loss = self.vq(x, n_q=n_q)
        bw = torch.tensor(n_q * self.get_bandwidth_per_quantizer(sample_rate)).to(x)
        return quantized, codes, bw, torch.mean(commit_loss)

    def encode(self, x: torch.Tensor, sample_rate: int, bandwidth: tp.Optional[float] = None,
               st: tp.Optional[int] = None) -> torch.Tensor:
        n_q = self.get_num_quantizers_for_bandwidth(sample_rate, bandwidth)
        return self.vq.encode(x, n_q=n_q, st=st or 0)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        return self.vq.decode(codes)
