"""Model entry points around the quantizer.

The reference models are `encoder -> quantizer -> decoder` sandwiches:
  SoundStream.forward / encode / decode     academicodec/models/encodec/net3.py:38-61
  VQVAE.encode / forward                    academicodec/models/hificodec/vqvae.py:31-45
The conv nets on either side stay stock PyTorch (cuDNN); only the quantizer call in the middle
is this package's.  This module provides (a) `swap_quantizer`, which replaces the reference
quantizer of an already-built model in place, and (b) two thin wrappers with the reference's
entry-point signatures for callers that bring their own encoder / decoder modules.
"""
from __future__ import annotations

import math
import random
import typing as tp

import torch
from torch import nn

from .grvq import Quantizer
from .quantization import ResidualVectorQuantizer


def _convert_rvq(ref: nn.Module) -> ResidualVectorQuantizer:
    """Build our ResidualVectorQuantizer from a reference one (same hyper-parameters, same state)."""
    new = ResidualVectorQuantizer(dimension=ref.dimension, n_q=ref.n_q, bins=ref.bins, decay=ref.decay,
                                  kmeans_init=ref.kmeans_init, kmeans_iters=ref.kmeans_iters,
                                  threshold_ema_dead_code=ref.threshold_ema_dead_code)
    new.load_state_dict(ref.state_dict())
    dev = next(ref.buffers()).device
    return new.to(dev).train(ref.training)


def _convert_grvq(ref: nn.Module) -> Quantizer:
    new = Quantizer(ref.h)
    new.load_state_dict(ref.state_dict())
    dev = next(ref.parameters()).device
    return new.to(dev).train(ref.training)


def swap_quantizer(model: nn.Module, attr: str = "quantizer") -> nn.Module:
    """Replace `model.<attr>` (a reference ResidualVectorQuantizer or HiFi-Codec Quantizer) by
    the B200-native module carrying the same state.  Returns the model."""
    ref = getattr(model, attr)
    if hasattr(ref, "vq") and hasattr(ref, "bins"):
        setattr(model, attr, _convert_rvq(ref))
    elif hasattr(ref, "quantizer_modules") and hasattr(ref, "h"):
        setattr(model, attr, _convert_grvq(ref))
    else:
        raise TypeError(f"{type(ref).__name__} is neither an RVQ nor a GRVQ quantizer")
    return model


class SoundStreamCodec(nn.Module):
    """`SoundStream`-shaped wrapper (net3.py:12-61) around caller-supplied encoder / decoder
    modules (e.g. the reference SEANetEncoder / SEANetDecoder) and the B200 quantizer."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module, D: int = 512,
                 target_bandwidths: tp.Sequence[float] = (7.5, 15),
                 ratios: tp.Sequence[int] = (8, 5, 4, 2), sample_rate: int = 24000, bins: int = 1024):
        super().__init__()
        self.hop_length = int(math.prod(ratios))
        self.encoder = encoder
        # n_q and frame_rate exactly as net3.py:25-26
        n_q = int(1000 * target_bandwidths[-1] // (math.ceil(sample_rate / self.hop_length) * 10))
        self.frame_rate = math.ceil(sample_rate / math.prod(ratios))
        self.bits_per_codebook = int(math.log2(bins))
        self.target_bandwidths = list(target_bandwidths)
        self.quantizer = ResidualVectorQuantizer(dimension=D, n_q=n_q, bins=bins)
        self.decoder = decoder

    def get_last_layer(self):
        return self.decoder.layers[-1].weight

    def forward(self, x: torch.Tensor) -> tp.Tuple[torch.Tensor, torch.Tensor, None]:
        e = self.encoder(x)
        bw = self.target_bandwidths[random.randint(0, len(self.target_bandwidths) - 1)]
        quantized, codes, bandwidth, commit_loss = self.quantizer(e, self.frame_rate, bw)
        return self.decoder(quantized), commit_loss, None

    def encode(self, x: torch.Tensor, target_bw: tp.Optional[float] = None,
               st: tp.Optional[int] = None) -> torch.Tensor:
        e = self.encoder(x)
        bw = self.target_bandwidths[-1] if target_bw is None else target_bw
        return self.quantizer.encode(e, self.frame_rate, bw, st or 0)

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        return self.decoder(self.quantizer.decode(codes))


class HiFiCodec(nn.Module):
    """`VQVAE`-shaped wrapper (vqvae.py:12-45) around caller-supplied HiFi-Codec `Encoder` /
    `Generator` modules and the B200 GRVQ quantizer (no checkpoint file needed to construct)."""

    def __init__(self, h, encoder: nn.Module, generator: nn.Module):
        super().__init__()
        self.h = h
        self.quantizer = Quantizer(h)
        self.generator = generator
        self.encoder = encoder

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """codes [B, T, 2G] -> waveform (vqvae.py:31-35)."""
        return self.generator(self.quantizer.embed(x))

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """waveform [B, L] -> codes [B, T, 2G] (vqvae.py:37-45)."""
        batch_size = x.size(0)
        if len(x.shape) == 3 and x.shape[-1] == 1:
            x = x.squeeze(-1)
        c = self.encoder(x.unsqueeze(1))
        # the reference runs the full forward and drops quantized / loss (vqvae.py:42);
        # the codes-only search gives the same indices on the tensor-core kernel
        codes = self.quantizer.encode(c)
        codes = [code.reshape(batch_size, -1) for code in codes]
        return torch.stack(codes, -1)
