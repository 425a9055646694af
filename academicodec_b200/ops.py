"""Functional layer over the C ABI: torch tensors in, torch tensors out.

PyTorch is only the owner of device memory and streams here; every arithmetic step runs in
the hand-written sm_100a kernels of libacq_b200.so, on the caller's current CUDA stream.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACQ_IMPL_AUTO, ACQ_LOSS_RAW, ACQ_STE  # noqa: F401  (re-exported)


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (academicodec_b200 has no CPU path); got {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


def _check_tables(codebooks: Sequence[torch.Tensor], n: int, k: int, dg: int, device) -> List[torch.Tensor]:
    if len(codebooks) != n:
        raise ValueError(f"expected {n} codebooks, got {len(codebooks)}")
    if n > _lib.ACQ_MAX_TABLE:
        raise ValueError(f"stages*groups = {n} exceeds {_lib.ACQ_MAX_TABLE}")
    out = []
    for i, cb in enumerate(codebooks):
        _require_cuda_f32(cb, f"codebook[{i}]")
        if cb.device != device:
            raise RuntimeError(f"codebook[{i}] on {cb.device}, latents on {device}")
        if tuple(cb.shape) != (k, dg):
            raise ValueError(f"codebook[{i}] has shape {tuple(cb.shape)}, expected {(k, dg)}")
        out.append(cb if cb.is_contiguous() else cb.contiguous())
    return out


def codebook_half_norms(codebooks: Sequence[torch.Tensor]) -> torch.Tensor:
    """0.5*||e_k||^2 for every table -> [n_tables, K] fp32 (include/acq_b200.h)."""
    k, dg = codebooks[0].shape
    dev = codebooks[0].device
    cbs = _check_tables(codebooks, len(codebooks), k, dg, dev)
    out = torch.empty((len(cbs), k), dtype=torch.float32, device=dev)
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().acq_codebook_half_norms(tab, len(cbs), k, dg, out.data_ptr(),
                                                        _stream(dev)), "acq_codebook_half_norms")
    del keep
    return out


def rvq_search(x: torch.Tensor, codebooks: Sequence[torch.Tensor], stages: int, groups: int = 1,
               half_norms: Optional[torch.Tensor] = None, flags: int = 0, impl: int = ACQ_IMPL_AUTO,
               want_quantized: bool = False, want_residual: bool = False,
               want_sqerr: bool = False
               ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused residual nearest-codeword search (acq_rvq_search).

    x [B, D, T] fp32; codebooks stage-major list of `stages*groups` tensors [K, D/groups].
    -> codes [stages*groups, B*T] int64, quantized [B,D,T] | None, residual | None,
       sqerr [stages] fp64 | None
    """
    _require_cuda_f32(x, "x")
    if x.dim() != 3:
        raise ValueError(f"x must be [B, D, T], got {tuple(x.shape)}")
    b, d, t = x.shape
    if d % groups:
        raise ValueError(f"D={d} not divisible by groups={groups}")
    k = codebooks[0].shape[0]
    cbs = _check_tables(codebooks, stages * groups, k, d // groups, x.device)
    x = x.contiguous()
    if half_norms is None:
        half_norms = codebook_half_norms(cbs)
    dev = x.device
    codes = torch.empty((stages * groups, b * t), dtype=torch.int64, device=dev)
    quantized = torch.empty_like(x) if want_quantized else None
    residual = torch.empty_like(x) if want_residual else None
    sqerr = torch.zeros((stages,), dtype=torch.float64, device=dev) if want_sqerr else None
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        rc = _lib.load().acq_rvq_search(
            x.data_ptr(), tab, half_norms.data_ptr(), stages, groups, k, d, b, t, flags, impl,
            codes.data_ptr(), quantized.data_ptr() if want_quantized else None,
            residual.data_ptr() if want_residual else None,
            sqerr.data_ptr() if want_sqerr else None, _stream(dev))
    _lib.check(rc, "acq_rvq_search")
    del keep
    return codes, quantized, residual, sqerr


def vq_decode(codes: torch.Tensor, stride_table: int, stride_frame: int,
              codebooks: Sequence[torch.Tensor], stages: int, groups: int, batch: int, frames: int,
              check: bool = True) -> torch.Tensor:
    """Codebook gather-accumulate (acq_vq_decode) -> [B, D, T] fp32.

    With check=True an out-of-range code raises IndexError (what F.embedding does in the
    reference); this reads one flag back from the device, i.e. synchronises."""
    if not codes.is_cuda:
        raise RuntimeError("codes must be a CUDA tensor (academicodec_b200 has no CPU path)")
    if codes.dtype != torch.int64:
        raise TypeError(f"codes must be int64, got {codes.dtype}")
    k, dg = codebooks[0].shape
    d = dg * groups
    dev = codes.device
    cbs = _check_tables(codebooks, stages * groups, k, dg, dev)
    codes = codes.contiguous()
    out = torch.empty((batch, d, frames), dtype=torch.float32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev) if check else None
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        rc = _lib.load().acq_vq_decode(codes.data_ptr(), stride_table, stride_frame, tab, stages,
                                       groups, k, d, batch, frames, out.data_ptr(),
                                       status.data_ptr() if check else None, _stream(dev))
    _lib.check(rc, "acq_vq_decode")
    del keep
    if check and int(status.item()) != 0:
        raise IndexError("index out of range in codes (valid range [0, %d))" % k)
    return out


def ema_stats(x: torch.Tensor, codes: torch.Tensor, codebooks: Sequence[torch.Tensor],
              flags: int = ACQ_STE) -> torch.Tensor:
    """Cluster sums and counts for every stage (acq_ema_stats) -> flat fp32
    [S*K*D sums | S*K counts], ready for one all-reduce."""
    _require_cuda_f32(x, "x")
    b, d, t = x.shape
    s = len(codebooks)
    k = codebooks[0].shape[0]
    cbs = _check_tables(codebooks, s, k, d, x.device)
    x = x.contiguous()
    codes = codes.contiguous()
    if tuple(codes.shape) != (s, b * t) and tuple(codes.shape) != (s, b, t):
        raise ValueError(f"codes shape {tuple(codes.shape)} does not match [S={s}, B*T={b * t}]")
    stats = torch.zeros((s * k * (d + 1),), dtype=torch.float32, device=x.device)
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(x.device):
        rc = _lib.load().acq_ema_stats(x.data_ptr(), codes.data_ptr(), tab, s, k, d, b, t, flags,
                                       stats.data_ptr(), _stream(x.device))
    _lib.check(rc, "acq_ema_stats")
    del keep
    return stats


def ema_apply(stats: torch.Tensor, embed: Sequence[torch.Tensor], embed_avg: Sequence[torch.Tensor],
              cluster_size: Sequence[torch.Tensor], decay: float, epsilon: float) -> None:
    """In-place EMA / Laplace / normalise on the module buffers (acq_ema_apply)."""
    s = len(embed)
    k, d = embed[0].shape
    for group in (embed, embed_avg, cluster_size):
        for tns in group:
            _require_cuda_f32(tns, "EMA buffer")
            if not tns.is_contiguous():
                raise RuntimeError("EMA buffers must be contiguous")
    t1, k1 = _lib.ptr_table(list(embed))
    t2, k2 = _lib.ptr_table(list(embed_avg))
    t3, k3 = _lib.ptr_table(list(cluster_size))
    dev = stats.device
    with torch.cuda.device(dev):
        rc = _lib.load().acq_ema_apply(stats.data_ptr(), t1, t2, t3, s, k, d, float(decay),
                                       float(epsilon), _stream(dev))
    _lib.check(rc, "acq_ema_apply")
    del k1, k2, k3


class HostPipeline:
    """Host-buffer encode/decode (acq_pipeline_*): pinned host tensors in, pinned host tensors
    out; H2D copy, kernels and D2H copy of consecutive chunks overlap on a ring of streams."""

    def __init__(self, device: int = 0, chunk_bytes: int = 64 << 20):
        import ctypes
        self._h = ctypes.c_void_p()
        _lib.check(_lib.load().acq_pipeline_create(ctypes.byref(self._h), int(device), int(chunk_bytes)),
                   "acq_pipeline_create")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.load().acq_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launches(self) -> int:
        return int(_lib.load().acq_pipeline_last_launches(self._h))

    def rvq_encode(self, x_host: torch.Tensor, codebooks, stages: int, groups: int,
                   half_norms: torch.Tensor, flags: int = 0, impl: int = ACQ_IMPL_AUTO,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous float32 CPU tensor")
        b, d, t = x_host.shape
        k = codebooks[0].shape[0]
        if out is None:
            out = torch.empty((stages * groups, b * t), dtype=torch.int64, pin_memory=True)
        tab, keep = _lib.ptr_table(list(codebooks))
        rc = _lib.load().acq_rvq_encode_host(self._h, x_host.data_ptr(), tab, half_norms.data_ptr(),
                                             stages, groups, k, d, b, t, flags, impl, out.data_ptr())
        _lib.check(rc, "acq_rvq_encode_host")
        del keep
        return out

    def vq_decode(self, codes_host: torch.Tensor, stride_table: int, stride_frame: int, codebooks,
                  stages: int, groups: int, batch: int, frames: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if codes_host.is_cuda or codes_host.dtype != torch.int64 or not codes_host.is_contiguous():
            raise ValueError("codes_host must be a contiguous int64 CPU tensor")
        k, dg = codebooks[0].shape
        d = dg * groups
        if out is None:
            out = torch.empty((batch, d, frames), dtype=torch.float32, pin_memory=True)
        tab, keep = _lib.ptr_table(list(codebooks))
        rc = _lib.load().acq_vq_decode_host(self._h, codes_host.data_ptr(), stride_table, stride_frame,
                                            tab, stages, groups, k, d, batch, frames, out.data_ptr())
        _lib.check(rc, "acq_vq_decode_host")
        del keep
        return out
