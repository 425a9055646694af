"""CUDA-graph replay of the codec calls at a fixed shape (small-batch latency).

At the reference's recipe batch (16 clips x 1 s = 1 600 frames) one encode is a ~0.15 ms kernel
behind tens of microseconds of Python / ctypes / launch overhead per call, and a decode is a 25 us
kernel behind the same overhead.  Every C-ABI entry point used here is capturable -- no host
synchronisation, no allocation outside torch's allocator, device pointers passed by value -- so a
streaming caller can record `encode` and `decode` once and replay them.

    g = GraphedCodec(quantizer, example_x, frame_rate=100, bandwidth=6.0)
    codes = g.encode(x)          # x: same shape / dtype as example_x; returns the graph's static buffer
    latents = g.decode(codes)

The graphs read the codebooks, their half norms and the tensor-core operand pack through the pointers
captured at record time.  The object keeps those derived tensors alive, so a replay never touches freed
memory, but after the codebooks change (an EMA update, load_state_dict) it would search against the OLD
norms / pack: call `GraphedCodec.record()` again after training steps.
"""
from __future__ import annotations

import typing as tp

import torch

from . import ops


class GraphedCodec:
    def __init__(self, quantizer, example_x: torch.Tensor, frame_rate: int,
                 bandwidth: tp.Optional[float] = None):
        if not example_x.is_cuda:
            raise RuntimeError("GraphedCodec needs CUDA tensors (academicodec_b200 has no CPU path)")
        self.q = quantizer
        self.frame_rate = frame_rate
        self.bandwidth = bandwidth
        self._x = example_x.detach().clone().contiguous()
        self.record()

    def record(self) -> None:
        """(Re-)record both graphs against the quantizer's current codebooks."""
        q, dev = self.q, self._x.device
        b, d, t = self._x.shape
        # warm up and capture on the same stream: the tensor-core workspace is cached per stream, and a
        # capture on another stream would allocate and zero-fill a fresh one inside the graph
        side = self._stream = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                          # warm-up: norms / operand pack / workspace caches
                codes = q.encode(self._x, self.frame_rate, self.bandwidth)
            # the captured launches hold raw pointers into the module's cached derived tables: own them
            self._tables = (getattr(q.vq, "_norm_cache", None), getattr(q.vq, "_tc_cache", None))
            s = codes.shape[0]
            self._embeds = [layer._codebook.embed for layer in q.vq.layers[:s]]
            self._codes_in = codes.clone()
            self._out = torch.empty((b, d, t), dtype=torch.float32, device=dev)
            ops.vq_decode(self._codes_in, b * t, 1, self._embeds, s, 1, b, t, check=False, out=self._out)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._g_enc = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_enc, stream=side), torch.no_grad():
            self._codes = q.encode(self._x, self.frame_rate, self.bandwidth)
        self._g_dec = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_dec, stream=side), torch.no_grad():
            ops.vq_decode(self._codes_in, b * t, 1, self._embeds, s, 1, b, t, check=False, out=self._out)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, D, T] (the recorded shape) -> codes [n_q', B, T] int64 (static buffer of the graph)."""
        if x.shape != self._x.shape or x.dtype != self._x.dtype:
            raise ValueError(f"graph was recorded for {tuple(self._x.shape)} {self._x.dtype}")
        self._x.copy_(x, non_blocking=True)
        self._g_enc.replay()
        return self._codes

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [n_q', B, T] int64 -> latents [B, D, T] (static buffer).  Codes are not range-checked
        (the check reads a flag back, i.e. synchronises); out-of-range codes contribute zero."""
        if codes.shape != self._codes_in.shape or codes.dtype != torch.int64:
            raise ValueError(f"graph was recorded for codes {tuple(self._codes_in.shape)} int64")
        if codes.data_ptr() != self._codes_in.data_ptr():
            self._codes_in.copy_(codes, non_blocking=True)
        self._g_dec.replay()
        return self._out
