"""Functional layer over the C ABI: torch tensors in, torch tensors out.

PyTorch is only the owner of device memory and streams here; every arithmetic step runs in
the hand-written sm_100a kernels of libacq_b200.so, on the caller's current CUDA stream.
"""
from __future__ import annotations

import os
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


def tc_supported(k: int, d: int, groups: int = 1) -> bool:
    """Shapes the tcgen05 kernel accepts (include/acq_b200.h)."""
    dg = d // max(groups, 1)
    return (groups <= 4 and d % groups == 0 and k % 256 == 0 and k <= 1024
            and dg % 64 == 0 and dg <= 512)


def tc_pack_table_bytes(k: int, dg: int) -> int:
    return int(_lib.load().acq_tc_pack_bytes(1, k, dg))


def tc_pack_codebooks(codebooks: Sequence[torch.Tensor]) -> torch.Tensor:
    """Tensor-core operand images + scaled half norms for the given tables (acq_tc_pack_codebooks).
    Returns an opaque uint8 device buffer of len(codebooks) equal-sized records, so
    `pack[i * tc_pack_table_bytes(k, dg):]` is the pack of tables i.. ; rebuild it whenever a
    codebook changes."""
    k, dg = codebooks[0].shape
    dev = codebooks[0].device
    cbs = _check_tables(codebooks, len(codebooks), k, dg, dev)
    lib = _lib.load()
    nbytes = int(lib.acq_tc_pack_bytes(len(cbs), k, dg))
    pack = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        _lib.check(lib.acq_tc_pack_codebooks(tab, len(cbs), k, dg, pack.data_ptr(), _stream(dev)),
                   "acq_tc_pack_codebooks")
    del keep
    return pack


_workspaces = {}


def tc_workspace(d: int, device: torch.device) -> torch.Tensor:
    """Per-(device, stream, D) scratch for the tensor-core kernel (residual rows of the tiles in
    flight).  Calls on the same stream are serialised, so they may share it."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, d)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros((int(_lib.load().acq_tc_workspace_bytes(d)),), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def debug_tc_scores(x: torch.Tensor, codebook: torch.Tensor):
    """Test hook: (scores [B*T, K] = x.e_k - 0.5||e_k||^2 as the tensor-core path computed them,
    codes [B*T])."""
    b, d, t = x.shape
    k = codebook.shape[0]
    pack = tc_pack_codebooks([codebook])
    ws = tc_workspace(d, x.device)
    scores = torch.zeros((b * t, k), dtype=torch.float32, device=x.device)
    codes = torch.empty((b * t,), dtype=torch.int64, device=x.device)
    tab, keep = _lib.ptr_table([codebook])
    with torch.cuda.device(x.device):
        rc = _lib.load().acq_debug_tc_scores(x.contiguous().data_ptr(), tab, pack.data_ptr(), ws.data_ptr(),
                                             k, d, b, t, scores.data_ptr(), codes.data_ptr(),
                                             _stream(x.device))
    _lib.check(rc, "acq_debug_tc_scores")
    del keep
    # the dump is scaled by the pack's power-of-two codebook scale cs; a table's record is
    # [images | norms | tail: cs, ... (256 B) | bias images: K/256 x 16 KiB]
    tail = pack.numel() - 256 - (k // 256) * 16384
    cs = pack[tail: tail + 4].view(torch.float32)
    return scores / cs, codes


def rvq_search(x: torch.Tensor, codebooks: Sequence[torch.Tensor], stages: int, groups: int = 1,
               half_norms: Optional[torch.Tensor] = None, flags: int = 0, impl: int = ACQ_IMPL_AUTO,
               want_quantized: bool = False, want_residual: bool = False,
               want_sqerr: bool = False, tc_pack: Optional[torch.Tensor] = None,
               codes_out: Optional[torch.Tensor] = None
               ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused residual nearest-codeword search (acq_rvq_search).

    x [B, D, T] fp32; codebooks stage-major list of `stages*groups` tensors [K, D/groups].
    `tc_pack` (from tc_pack_codebooks for exactly these tables) enables the tcgen05 kernel for
    codes-only calls; without it, or when quantized/residual/sqerr are requested, the fused SIMT
    kernel runs.
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
    dev = x.device
    codes_only = not (want_quantized or want_residual or want_sqerr)
    # the tensor-core kernel writes codes only; under AUTO the library follows it with the replay pass
    # when other outputs are requested, ACQ_IMPL_TC insists on a codes-only call
    use_tc = (tc_pack is not None and impl != _lib.ACQ_IMPL_SIMT and tc_supported(k, d, groups)
              and (codes_only or impl == _lib.ACQ_IMPL_AUTO))
    if impl == _lib.ACQ_IMPL_TC and not use_tc:
        raise ValueError("tensor-core search needs tc_pack, a supported shape and a codes-only call")
    workspace = tc_workspace(d, dev) if use_tc else None
    if half_norms is None and not use_tc:
        half_norms = codebook_half_norms(cbs)
    if codes_out is not None:
        if codes_out.dtype != torch.int64 or codes_out.numel() != stages * groups * b * t \
                or not codes_out.is_contiguous() or codes_out.device != dev:
            raise ValueError("codes_out must be a contiguous int64 tensor of S*G*B*T elements on x's device")
        codes = codes_out.view(stages * groups, b * t)
    else:
        codes = torch.empty((stages * groups, b * t), dtype=torch.int64, device=dev)
    quantized = torch.empty_like(x) if want_quantized else None
    residual = torch.empty_like(x) if want_residual else None
    sqerr = torch.zeros((stages,), dtype=torch.float64, device=dev) if want_sqerr else None
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        rc = _lib.load().acq_rvq_search(
            x.data_ptr(), tab, half_norms.data_ptr() if half_norms is not None else None,
            tc_pack.data_ptr() if use_tc else None, workspace.data_ptr() if use_tc else None,
            stages, groups, k, d, b, t, flags, impl,
            codes.data_ptr(), quantized.data_ptr() if want_quantized else None,
            residual.data_ptr() if want_residual else None,
            sqerr.data_ptr() if want_sqerr else None, _stream(dev))
    _lib.check(rc, "acq_rvq_search")
    del keep
    return codes, quantized, residual, sqerr


class _DeferredCodeCheck:
    """Out-of-range codes without a host synchronisation.

    The reference's F.embedding raises IndexError on the CPU; on a CUDA device the same call trips a
    device-side assert that surfaces at a later synchronisation point.  The module-level decode paths
    behave like the latter: the kernel raises a device flag, the flag is copied to pinned host memory
    behind the kernel, and the *next* decode on that device (or `check_codes_now`) raises IndexError once
    the copy has landed.  Nothing blocks, and the calls stay CUDA-graph capturable (no check is recorded
    while a stream is capturing)."""

    def __init__(self, device: torch.device):
        self.flag = torch.zeros((1,), dtype=torch.int32, device=device)
        self.host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event()
        self.pending = False
        self.k = 0

    def poll(self, block: bool = False) -> None:
        if not self.pending:
            return
        if block:
            self.event.synchronize()
        elif not self.event.query():
            return
        self.pending = False
        if int(self.host[0]) != 0:
            self.host.zero_()
            self.flag.zero_()
            raise IndexError("index out of range in codes of an earlier decode (valid range [0, %d))" % self.k)

    def arm(self, k: int) -> None:
        self.k = k
        self.host.copy_(self.flag, non_blocking=True)
        self.event.record()
        self.pending = True


_deferred_checks = {}


def _deferred(device: torch.device) -> _DeferredCodeCheck:
    c = _deferred_checks.get(device.index)
    if c is None:
        c = _deferred_checks[device.index] = _DeferredCodeCheck(device)
    return c


def check_codes_now(device=None) -> None:
    """Wait for outstanding deferred code-range checks and raise IndexError if one failed."""
    for idx, c in list(_deferred_checks.items()):
        if device is None or torch.device(device).index in (None, idx):
            c.poll(block=True)


# ACQ_DEBUG=1: module-level decodes check their codes synchronously (IndexError at the call site)
_SYNC_CHECK = os.environ.get("ACQ_DEBUG", "0") not in ("", "0")


def vq_decode(codes: torch.Tensor, stride_table: int, stride_frame: int,
              codebooks: Sequence[torch.Tensor], stages: int, groups: int, batch: int, frames: int,
              check="deferred", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Codebook gather-accumulate (acq_vq_decode) -> [B, D, T] fp32.

    check=True: an out-of-range code raises IndexError at the call (what F.embedding does on the CPU);
    this reads one flag back from the device, i.e. synchronises.  check="deferred" (default, what the
    modules use): no synchronisation, the error is raised by a later decode / `check_codes_now`
    (_DeferredCodeCheck).  check=False: no check; out-of-range codes contribute zeros."""
    if not codes.is_cuda:
        raise RuntimeError("codes must be a CUDA tensor (academicodec_b200 has no CPU path)")
    if codes.dtype != torch.int64:
        raise TypeError(f"codes must be int64, got {codes.dtype}")
    k, dg = codebooks[0].shape
    d = dg * groups
    dev = codes.device
    cbs = _check_tables(codebooks, stages * groups, k, dg, dev)
    codes = codes.contiguous()
    if out is not None:
        if out.dtype != torch.float32 or tuple(out.shape) != (batch, d, frames) \
                or not out.is_contiguous() or out.device != dev:
            raise ValueError("out must be a contiguous float32 [B, D, T] tensor on the codes' device")
    else:
        out = torch.empty((batch, d, frames), dtype=torch.float32, device=dev)
    if check == "deferred" and _SYNC_CHECK:
        check = True
    deferred = None
    if check == "deferred":
        if torch.cuda.is_current_stream_capturing():
            check = False
        else:
            deferred = _deferred(dev)
            deferred.poll()
    status = torch.zeros((1,), dtype=torch.int32, device=dev) if check is True else (
        deferred.flag if deferred is not None else None)
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        rc = _lib.load().acq_vq_decode(codes.data_ptr(), stride_table, stride_frame, tab, stages,
                                       groups, k, d, batch, frames, out.data_ptr(),
                                       status.data_ptr() if status is not None else None, _stream(dev))
        _lib.check(rc, "acq_vq_decode")
        if deferred is not None:
            deferred.arm(k)
    del keep
    if check is True and int(status.item()) != 0:
        raise IndexError("index out of range in codes (valid range [0, %d))" % k)
    return out


def rvq_replay(x: torch.Tensor, codes: torch.Tensor, codebooks: Sequence[torch.Tensor], stages: int,
               groups: int = 1, flags: int = 0, want_quantized: bool = True, want_residual: bool = False,
               want_sqerr: bool = False, want_stats: bool = False):
    """Everything forward() returns besides the codes, from x and the codes (acq_rvq_replay).
    -> (quantized | None, residual | None, sqerr [S] fp64 | None, stats flat fp32 | None)"""
    _require_cuda_f32(x, "x")
    b, d, t = x.shape
    k = codebooks[0].shape[0]
    cbs = _check_tables(codebooks, stages * groups, k, d // groups, x.device)
    x = x.contiguous()
    codes = codes.contiguous()
    if codes.numel() != stages * groups * b * t or codes.dtype != torch.int64:
        raise ValueError("codes must be int64 with S*G*B*T elements")
    dev = x.device
    quantized = torch.empty_like(x) if want_quantized else None
    residual = torch.empty_like(x) if want_residual else None
    sqerr = torch.zeros((stages,), dtype=torch.float64, device=dev) if want_sqerr else None
    stats = torch.zeros((stages * k * (d + 1),), dtype=torch.float32, device=dev) if want_stats else None
    tab, keep = _lib.ptr_table(cbs)
    with torch.cuda.device(dev):
        rc = _lib.load().acq_rvq_replay(
            x.data_ptr(), codes.data_ptr(), tab, stages, groups, k, d, b, t, flags,
            quantized.data_ptr() if want_quantized else None,
            residual.data_ptr() if want_residual else None,
            sqerr.data_ptr() if want_sqerr else None,
            stats.data_ptr() if want_stats else None, _stream(dev))
    _lib.check(rc, "acq_rvq_replay")
    del keep
    return quantized, residual, sqerr, stats


def grvq_backward(x: torch.Tensor, codes: torch.Tensor, codebooks: Sequence[torch.Tensor], stages: int, groups: int,
                  g_quantized, g_losses, lam_cb: float, lam_commit: float, want_grad_x: bool = True,
                  want_grad_cb: Sequence[bool] | None = None):
    """Gradients of the group-residual VQ training forward in one kernel (acq_grvq_backward).
    -> (grad_x [B, D, T] | None, [grad of codebook i [K, D/G] | None, ...])"""
    _require_cuda_f32(x, "x")
    b, d, t = x.shape
    k = codebooks[0].shape[0]
    n_tab = stages * groups
    cbs = _check_tables(codebooks, n_tab, k, d // groups, x.device)
    x = x.contiguous()
    codes = codes.contiguous()
    if codes.numel() != n_tab * b * t or codes.dtype != torch.int64:
        raise ValueError("codes must be int64 with S*G*B*T elements")
    dev = x.device
    want = list(want_grad_cb) if want_grad_cb is not None else [True] * n_tab
    grad_x = torch.empty_like(x) if want_grad_x else None
    grads = [torch.zeros_like(cbs[i]) if want[i] else None for i in range(n_tab)]
    if g_quantized is not None:
        _require_cuda_f32(g_quantized, "g_quantized")
        g_quantized = g_quantized.contiguous()
    if g_losses is not None:
        g_losses = g_losses.to(torch.float32).contiguous()
    tab, keep = _lib.ptr_table(cbs)
    import ctypes as _ct
    garr = (_ct.c_void_p * n_tab)(*[g.data_ptr() if g is not None else None for g in grads])
    with torch.cuda.device(dev):
        rc = _lib.load().acq_grvq_backward(
            x.data_ptr(), codes.data_ptr(), tab, stages, groups, k, d, b, t,
            g_quantized.data_ptr() if g_quantized is not None else None,
            g_losses.data_ptr() if g_losses is not None else None,
            float(lam_cb), float(lam_commit), grad_x.data_ptr() if want_grad_x else None,
            _ct.cast(garr, _ct.POINTER(_ct.c_void_p)), _stream(dev))
    _lib.check(rc, "acq_grvq_backward")
    del keep, garr
    return grad_x, grads


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


def pack_codes(values: torch.Tensor, bits: int, check: bool = True) -> torch.Tensor:
    """Pack int64 codes into the reference's `bits`-wide little-endian bit stream
    (acq_pack_codes; reference binary.py BitPacker) -> uint8 tensor of ceil(n*bits/8) bytes."""
    if not values.is_cuda or values.dtype != torch.int64:
        raise TypeError("values must be a CUDA int64 tensor")
    v = values.contiguous().view(-1)
    n = v.numel()
    lib = _lib.load()
    out = torch.empty((int(lib.acq_packed_bytes(n, bits)),), dtype=torch.uint8, device=v.device)
    status = torch.zeros((1,), dtype=torch.int32, device=v.device) if check else None
    with torch.cuda.device(v.device):
        rc = lib.acq_pack_codes(v.data_ptr(), n, bits, out.data_ptr(),
                                status.data_ptr() if check else None, _stream(v.device))
    _lib.check(rc, "acq_pack_codes")
    if check and int(status.item()) != 0:
        raise ValueError(f"a value does not fit in {bits} bits")
    return out


def unpack_codes(packed: torch.Tensor, n: int, bits: int) -> torch.Tensor:
    """Inverse of pack_codes (acq_unpack_codes; reference binary.py BitUnpacker) -> int64 [n]."""
    if not packed.is_cuda or packed.dtype != torch.uint8:
        raise TypeError("packed must be a CUDA uint8 tensor")
    lib = _lib.load()
    if packed.numel() < int(lib.acq_packed_bytes(n, bits)):
        raise ValueError("packed buffer too short")
    out = torch.empty((n,), dtype=torch.int64, device=packed.device)
    with torch.cuda.device(packed.device):
        rc = lib.acq_unpack_codes(packed.contiguous().data_ptr(), n, bits, out.data_ptr(), _stream(packed.device))
    _lib.check(rc, "acq_unpack_codes")
    return out


class HostPipeline:
    """Host-buffer encode/decode (acq_pipeline_*): pinned host tensors in, pinned host tensors
    out; H2D copy, kernels and D2H copy of consecutive chunks overlap on a ring of streams."""

    def __init__(self, device: int = 0, chunk_bytes: int = 128 << 20):
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

    def _prepare(self, *tables) -> None:
        """The pipeline's ring streams are independent of torch's: the tables this call reads
        (codebooks, half norms, tensor-core pack) must live on the pipeline's device, and whatever
        torch's current stream was given so far -- the kernels that produce those tables, an EMA
        update -- must complete before the ring streams touch them (acq_pipeline_wait_stream)."""
        for t in tables:
            if t is None:
                continue
            for u in (t if isinstance(t, (list, tuple)) else (t,)):
                if not u.is_cuda or u.device.index != self.device:
                    raise RuntimeError(f"HostPipeline on cuda:{self.device} got a table on {u.device}")
        dev = torch.device("cuda", self.device)
        _lib.check(_lib.load().acq_pipeline_wait_stream(self._h, _stream(dev)), "acq_pipeline_wait_stream")

    def rvq_encode(self, x_host: torch.Tensor, codebooks, stages: int, groups: int,
                   half_norms: torch.Tensor, flags: int = 0, impl: int = ACQ_IMPL_AUTO,
                   out: Optional[torch.Tensor] = None,
                   tc_pack: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous float32 CPU tensor")
        b, d, t = x_host.shape
        k = codebooks[0].shape[0]
        if out is None:
            out = torch.empty((stages * groups, b * t), dtype=torch.int64, pin_memory=True)
        self._prepare(list(codebooks), half_norms, tc_pack)
        tab, keep = _lib.ptr_table(list(codebooks))
        rc = _lib.load().acq_rvq_encode_host(self._h, x_host.data_ptr(), tab, half_norms.data_ptr(),
                                             tc_pack.data_ptr() if tc_pack is not None else None,
                                             stages, groups, k, d, b, t, flags, impl, out.data_ptr())
        _lib.check(rc, "acq_rvq_encode_host")
        del keep
        return out

    def rvq_codec(self, x_host: torch.Tensor, codebooks, stages: int, groups: int,
                  half_norms: torch.Tensor, flags: int = 0, impl: int = ACQ_IMPL_AUTO,
                  codes_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                  tc_pack: Optional[torch.Tensor] = None):
        """Encode then decode the same batch (acq_rvq_codec_host): host latents in, host codes
        [S*G, B*T] and host reconstructed latents [B, D, T] out; uploads and downloads of
        consecutive chunks overlap."""
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous float32 CPU tensor")
        b, d, t = x_host.shape
        k = codebooks[0].shape[0]
        if codes_out is None:
            codes_out = torch.empty((stages * groups, b * t), dtype=torch.int64, pin_memory=True)
        if out is None:
            out = torch.empty((b, d, t), dtype=torch.float32, pin_memory=True)
        self._prepare(list(codebooks), half_norms, tc_pack)
        tab, keep = _lib.ptr_table(list(codebooks))
        rc = _lib.load().acq_rvq_codec_host(self._h, x_host.data_ptr(), tab, half_norms.data_ptr(),
                                            tc_pack.data_ptr() if tc_pack is not None else None,
                                            stages, groups, k, d, b, t, flags, impl,
                                            codes_out.data_ptr(), out.data_ptr())
        _lib.check(rc, "acq_rvq_codec_host")
        del keep
        return codes_out, out

    def vq_decode(self, codes_host: torch.Tensor, stride_table: int, stride_frame: int, codebooks,
                  stages: int, groups: int, batch: int, frames: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if codes_host.is_cuda or codes_host.dtype != torch.int64 or not codes_host.is_contiguous():
            raise ValueError("codes_host must be a contiguous int64 CPU tensor")
        k, dg = codebooks[0].shape
        d = dg * groups
        if out is None:
            out = torch.empty((batch, d, frames), dtype=torch.float32, pin_memory=True)
        self._prepare(list(codebooks))
        tab, keep = _lib.ptr_table(list(codebooks))
        rc = _lib.load().acq_vq_decode_host(self._h, codes_host.data_ptr(), stride_table, stride_frame,
                                            tab, stages, groups, k, d, batch, frames, out.data_ptr())
        _lib.check(rc, "acq_vq_decode_host")
        del keep
        return out
