"""B200-native residual vector quantization modules.

Same class names, constructor arguments, method signatures, return tuples and state-dict keys
as the reference academicodec/quantization/core_vq.py (EuclideanCodebook :96-227,
VectorQuantization :230-315, ResidualVectorQuantization :318-370) so that checkpoints and
calling code carry over, but the modules hold no arithmetic: every forward/encode/decode is one
launch of the fused sm_100a kernels in libacq_b200.so (academicodec_b200/csrc), covering all
residual stages at once.  There is no CPU path.
"""
from __future__ import annotations

import typing as tp

import torch
from torch import nn

from .. import ops
from .distrib import all_reduce, broadcast_tensors, is_distributed


def _uniform_init(*shape: int) -> torch.Tensor:
    t = torch.empty(shape)
    nn.init.kaiming_uniform_(t)
    return t


@torch.no_grad()
def kmeans(data_bdt: torch.Tensor, means: torch.Tensor, num_iters: int = 10):
    """Lloyd iterations of the reference's `kmeans` (core_vq.py:72-93) on the device, starting from the
    given `means` [K, D] (the reference draws them from the data with torch.randperm, :61-69; the draw
    is the caller's).  Each iteration = nearest-codeword search (the fused search kernel) + per-cluster
    sums and counts (the EMA statistics kernel) + a divide; empty clusters keep their mean.
    data_bdt [B, D, T] -> (means [K, D], bins [K] = the last iteration's cluster populations)."""
    k, d = means.shape
    means = means.contiguous()
    counts = torch.zeros(k, device=data_bdt.device)
    for _ in range(num_iters):
        codes, _, _, _ = ops.rvq_search(data_bdt, [means], 1)
        stats = ops.ema_stats(data_bdt, codes, [means], flags=0)
        sums, counts = stats[:k * d].view(k, d), stats[k * d:]
        empty = counts == 0
        means = torch.where(empty[:, None], means, sums / counts.clamp(min=1)[:, None])
    return means, counts


class EuclideanCodebook(nn.Module):
    """One codebook with EMA k-means state.  Buffers (and hence state-dict keys) are the
    reference's: `inited [1]`, `cluster_size [K]`, `embed [K, D]`, `embed_avg [K, D]`
    (core_vq.py:134-137).  Tensors here are frames-last (`[..., D]`) as in the reference; the
    fused multi-stage path in ResidualVectorQuantization works on `[B, D, T]` directly."""

    def __init__(self, dim: int, codebook_size: int, kmeans_init: int = False,
                 kmeans_iters: int = 10, decay: float = 0.99, epsilon: float = 1e-5,
                 threshold_ema_dead_code: int = 2):
        super().__init__()
        self.decay = decay
        embed = torch.zeros(codebook_size, dim) if kmeans_init else _uniform_init(codebook_size, dim)
        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.epsilon = epsilon
        # Kept for signature compatibility.  Dead-code expiry in the reference rewrites `embed`,
        # which the EMA refresh overwrites a few lines later (core_vq.py:217 vs :224-225), so it
        # never changes the state; it is therefore not executed here ("expiry").
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.register_buffer("inited", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())
        self._kernel_writes = 0   # bumped when a kernel rewrites `embed` through its raw pointer
        self._inited_seen = False  # host-side memo of `inited` (reading the buffer synchronises)

    @property
    def is_inited(self) -> bool:
        if not self._inited_seen:
            self._inited_seen = bool(self.inited)
        return self._inited_seen

    def _load_from_state_dict(self, *args, **kwargs):
        self._inited_seen = False
        return super()._load_from_state_dict(*args, **kwargs)

    # -- cache key for data derived from `embed` (norms, packed tensor-core operands) ----------
    def cache_key(self):
        e = self.embed
        return (e.data_ptr(), e._version, self._kernel_writes, e.device)

    def invalidate_caches(self) -> None:
        """Call after writing `embed` behind autograd's back (`embed.data.copy_(w)`, a raw-pointer kernel
        write): such writes do not bump the tensor's version counter, so the half norms and tensor-core
        operand pack derived from the old values would otherwise be reused.  (`embed.copy_` under no_grad,
        `load_state_dict` and `.to()` are seen without it.)"""
        self._kernel_writes += 1

    # -- k-means initialisation on the first forward (core_vq.py:139-151, 72-93) ---------------
    @torch.no_grad()
    def init_embed_(self, data_bdt: torch.Tensor) -> None:
        if self.is_inited:
            return
        b, d, t = data_bdt.shape
        n, k = b * t, self.codebook_size
        if n >= k:
            pick = torch.randperm(n, device=data_bdt.device)[:k]
        else:
            pick = torch.randint(0, n, (k,), device=data_bdt.device)
        means = data_bdt.transpose(1, 2).reshape(n, d)[pick].contiguous()
        means, counts = kmeans(data_bdt, means, self.kmeans_iters)
        self.embed.data.copy_(means)
        self.embed_avg.data.copy_(means)
        self.cluster_size.data.copy_(counts)
        self.inited.data.fill_(1.0)
        self._inited_seen = True
        broadcast_tensors(self.buffers())
        self.invalidate_caches()       # `.data` writes do not bump the tensor version counter

    # -- frames-last helpers ----------------------------------------------------------------------
    @staticmethod
    def _as_bdt(x: torch.Tensor) -> torch.Tensor:
        flat = x.reshape(-1, x.shape[-1])
        return flat.t().contiguous().unsqueeze(0)          # [1, D, N]

    def quantize(self, x: torch.Tensor) -> torch.Tensor:
        """[N, D] -> [N] int64, nearest codeword, lowest index on ties (core_vq.py:175-180)."""
        codes, _, _, _ = ops.rvq_search(self._as_bdt(x), [self.embed], 1)
        return codes.view(-1)

    def dequantize(self, embed_ind: torch.Tensor) -> torch.Tensor:
        """[...] int64 -> [..., D] (core_vq.py:185-187)."""
        n = embed_ind.numel()
        out = ops.vq_decode(embed_ind.reshape(1, n), n, 1, [self.embed], 1, 1, 1, n)
        return out[0].t().reshape(*embed_ind.shape, self.embed.shape[1])

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.quantize(x.reshape(-1, x.shape[-1])).view(*x.shape[:-1])

    def decode(self, embed_ind: torch.Tensor) -> torch.Tensor:
        return self.dequantize(embed_ind)

    def forward(self, x: torch.Tensor):
        """[B, T, D] -> (quantize [B, T, D], embed_ind [B, T]); in training mode also performs the
        EMA update (core_vq.py:203-227)."""
        x_bdt = x.transpose(1, 2).contiguous() if x.dim() == 3 else self._as_bdt(x)
        self.init_embed_(x_bdt)
        b, d, t = x_bdt.shape
        codes, quant, _, _ = ops.rvq_search(x_bdt, [self.embed], 1, want_quantized=True)
        if self.training:
            ema_update_([self], x_bdt, codes, flags=0)
        quantize = quant.transpose(1, 2) if x.dim() == 3 else quant[0].t()
        return quantize.reshape(x.shape), codes.view(*x.shape[:-1])


def _peer_exchange(owner: "EuclideanCodebook", numel: int, device: torch.device):
    """The NVLink peer-memory exchange of a codebook stack (created collectively on first use, cached on the
    stack's first codebook); None = use NCCL (single process, ACQ_PEER_REDUCE=0, or no symmetric memory)."""
    cached = getattr(owner, "_peer_exchange", False)
    if cached is not False and (cached is None or (cached.numel >= numel and cached.buf.device == device)):
        return cached
    from .distrib import PeerExchange
    exch = PeerExchange.create(numel, device)
    object.__setattr__(owner, "_peer_exchange", exch)
    return exch


@torch.no_grad()
def ema_update_(codebooks: tp.Sequence[EuclideanCodebook], x_bdt: torch.Tensor, codes: torch.Tensor,
                flags: int, stats: tp.Optional[torch.Tensor] = None,
                all_codebooks: tp.Optional[tp.Sequence[EuclideanCodebook]] = None) -> None:
    """K3 statistics -> all-reduce(SUM) over ranks -> K4 apply, for every stage at once.

    The reference updates each rank from its local batch and relies on DDP re-broadcasting rank
    0's buffers (core_vq.py:218-225).  Here the statistics are made global with
    one all-reduce of a flat [S, K, D+1] buffer and every rank applies the identical update, so
    replicas stay bit-identical without any broadcast.

    `codebooks` are the stages this forward used; `all_codebooks` the whole stack.  The collective
    always has the size of the WHOLE stack: the number of stages used depends on the bandwidth, which
    the reference's SoundStream.forward draws per process (net3.py:41), so ranks may disagree on it and
    a size that followed n_q would dead-lock NCCL.  Stages a rank did not use contribute zeros; every
    rank then applies the update to every stage some rank used (acq_ema_apply skips stages whose
    reduced counts are all zero)."""
    embeds = [c.embed for c in codebooks]
    if stats is None:
        stats = ops.ema_stats(x_bdt, codes, embeds, flags=flags)
    if is_distributed():
        stack = list(all_codebooks) if all_codebooks is not None else list(codebooks)
        s, full = len(codebooks), len(stack)
        k, d = embeds[0].shape
        exch = _peer_exchange(stack[0], full * k * (d + 1), stats.device)
        if exch is not None:
            # NVLink peer-memory exchange: the statistics go into the symmetric buffer (whole-stack layout,
            # unused stages zero) and one kernel of this package sums them across the ranks in place
            buf = exch.buf
            if stats.data_ptr() != buf.data_ptr():
                if full > s:
                    buf.zero_()
                buf[:s * k * d].copy_(stats[:s * k * d])
                buf[full * k * d:full * k * d + s * k].copy_(stats[s * k * d:s * k * (d + 1)])
            stats = exch.all_reduce_()[:full * k * (d + 1)]
            codebooks = stack
            embeds = [c.embed for c in codebooks]
        else:
            if full > s:
                padded = torch.zeros(full * k * (d + 1), dtype=stats.dtype, device=stats.device)
                padded[:s * k * d] = stats[:s * k * d]
                padded[full * k * d:full * k * d + s * k] = stats[s * k * d:]
                stats, codebooks = padded, stack
                embeds = [c.embed for c in codebooks]
            all_reduce(stats)
    ops.ema_apply(stats, embeds, [c.embed_avg for c in codebooks],
                  [c.cluster_size for c in codebooks], codebooks[0].decay, codebooks[0].epsilon)
    for c in codebooks:
        c._kernel_writes += 1


class VectorQuantization(nn.Module):
    """One quantizer layer on `[B, D, T]` latents (reference core_vq.py:230-315)."""

    def __init__(self, dim: int, codebook_size: int, codebook_dim: tp.Optional[int] = None,
                 decay: float = 0.99, epsilon: float = 1e-5, kmeans_init: bool = True,
                 kmeans_iters: int = 50, threshold_ema_dead_code: int = 2,
                 commitment_weight: float = 1.0):
        super().__init__()
        _codebook_dim = codebook_dim if codebook_dim is not None else dim
        requires_projection = _codebook_dim != dim
        self.project_in = nn.Linear(dim, _codebook_dim) if requires_projection else nn.Identity()
        self.project_out = nn.Linear(_codebook_dim, dim) if requires_projection else nn.Identity()
        self.epsilon = epsilon
        self.commitment_weight = commitment_weight
        self._codebook = EuclideanCodebook(dim=_codebook_dim, codebook_size=codebook_size,
                                           kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                                           decay=decay, epsilon=epsilon,
                                           threshold_ema_dead_code=threshold_ema_dead_code)
        self.codebook_size = codebook_size

    @property
    def codebook(self):
        return self._codebook.embed

    @property
    def has_projection(self) -> bool:
        return not isinstance(self.project_in, nn.Identity)

    def _project_in(self, x_bdt: torch.Tensor) -> torch.Tensor:
        if not self.has_projection:
            return x_bdt
        return self.project_in(x_bdt.transpose(1, 2)).transpose(1, 2).contiguous()

    def _project_out(self, q_bdt: torch.Tensor) -> torch.Tensor:
        if not self.has_projection:
            return q_bdt
        return self.project_out(q_bdt.transpose(1, 2)).transpose(1, 2)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        x = self._project_in(x)
        b, _, t = x.shape
        codes, _, _, _ = ops.rvq_search(x, [self._codebook.embed], 1)
        return codes.view(b, t)

    def decode(self, embed_ind: torch.Tensor) -> torch.Tensor:
        b, t = embed_ind.shape
        q = ops.vq_decode(embed_ind, b * t, 1, [self._codebook.embed], 1, 1, b, t)
        return self._project_out(q)

    def forward(self, x: torch.Tensor):
        """-> (quantize [B, D, T], embed_ind [B, T], loss [1])."""
        xin = self._project_in(x)
        self._codebook.init_embed_(xin.detach())
        quantized, codes, losses = _stack_forward([self], xin, self.training,
                                                  _half_norms(self, [self]))
        return self._project_out(quantized), codes[0], losses[0]


class _ResidualSearchSTE(torch.autograd.Function):
    """Training-mode forward of a stack of layers with the reference's autograd contract:
    straight-through on the quantized sum, and the commitment loss reaching x through stage 0
    only (later residuals have zero Jacobian w.r.t. x because r - (r + (q - r).detach())
    cancels; SURVEY.md 8b 'Autograd')."""

    @staticmethod
    def forward(ctx, x, layers, half_norms, weights, tc_pack, stats_box):
        embeds = [layer._codebook.embed for layer in layers]
        b, d, t = x.shape
        if tc_pack is not None:
            # tensor-core search (codes), then one replay pass for the straight-through sum, the
            # commitment error and the EMA statistics
            codes, _, _, _ = ops.rvq_search(x, embeds, len(layers), half_norms=half_norms,
                                            flags=ops.ACQ_STE, tc_pack=tc_pack)
            quantized, _, sqerr, stats = ops.rvq_replay(x, codes, embeds, len(layers), 1, flags=ops.ACQ_STE,
                                                        want_sqerr=True, want_stats=True)
            stats_box.append(stats)
        else:
            codes, quantized, _, sqerr = ops.rvq_search(
                x, embeds, len(layers), half_norms=half_norms, flags=ops.ACQ_STE,
                want_quantized=True, want_sqerr=True)
        w = torch.tensor(weights, dtype=torch.float64, device=x.device)
        losses = (sqerr * w / float(x.numel())).to(x.dtype)
        if x.requires_grad:
            q0 = ops.vq_decode(codes[:1], b * t, 1, embeds[:1], 1, 1, b, t, check=False)
            ctx.save_for_backward(x - (x + (q0 - x)))        # x - q'_0
        ctx.scale = 2.0 * float(weights[0]) / float(x.numel())
        codes = codes.view(len(layers), b, t)
        ctx.mark_non_differentiable(codes)
        return quantized, codes, losses

    @staticmethod
    def backward(ctx, g_quantized, _g_codes, g_losses):
        (diff0,) = ctx.saved_tensors
        grad = g_quantized if g_quantized is not None else torch.zeros_like(diff0)
        if g_losses is not None:
            grad = grad + diff0 * (g_losses[0] * ctx.scale)
        return grad, None, None, None, None, None


def _half_norms(owner: nn.Module, layers) -> torch.Tensor:
    """0.5||e||^2 for `layers` -> [len(layers), K], cached on `owner` until a codebook changes."""
    key = tuple(layer._codebook.cache_key() for layer in layers)
    cached = getattr(owner, "_norm_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    hn = ops.codebook_half_norms([layer._codebook.embed for layer in layers])
    owner._norm_cache = (key, hn)
    return hn


def _tc_pack(owner: nn.Module, layers) -> tp.Optional[torch.Tensor]:
    """Tensor-core operand pack for `layers` (one record per layer, so `pack[i * rec:]` serves
    layers[i:]), cached on `owner` until a codebook changes; None if the shape is unsupported."""
    k, d = layers[0]._codebook.embed.shape
    if not ops.tc_supported(k, d):
        return None
    key = tuple(layer._codebook.cache_key() for layer in layers)
    cached = getattr(owner, "_tc_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    pack = ops.tc_pack_codebooks([layer._codebook.embed for layer in layers])
    owner._tc_cache = (key, pack)
    return pack


def _stack_forward(layers, x: torch.Tensor, training: bool, half_norms: torch.Tensor,
                   tc_pack: tp.Optional[torch.Tensor] = None, all_layers=None):
    """Fused forward over `layers` -> (quantized_out [B,D,T], codes [S,B,T], losses [S,1])."""
    s = len(layers)
    b, d, t = x.shape
    if training:
        weights = [float(layer.commitment_weight) for layer in layers]
        stats_box: tp.List[torch.Tensor] = []
        quantized, codes, losses = _ResidualSearchSTE.apply(x, layers, half_norms, weights, tc_pack,
                                                            stats_box)
        ema_update_([layer._codebook for layer in layers], x.detach(), codes.view(s, b * t),
                    flags=ops.ACQ_STE, stats=stats_box[0] if stats_box else None,
                    all_codebooks=[layer._codebook for layer in (all_layers or layers)])
        if not losses.requires_grad:
            losses = losses.clone().requires_grad_(True)   # reference: loss tensor requires grad
    else:
        embeds = [layer._codebook.embed for layer in layers]
        if tc_pack is not None:
            # tensor-core search for the codes, then the gather-accumulate kernel: decode(codes)
            # is bit-identical to the eval-mode quantized sum (0.0 + q_0 + q_1 + ...)
            codes, _, _, _ = ops.rvq_search(x, embeds, s, half_norms=half_norms, tc_pack=tc_pack)
            quantized = ops.vq_decode(codes, b * t, 1, embeds, s, 1, b, t, check=False)
        else:
            codes, quantized, _, _ = ops.rvq_search(x, embeds, s, half_norms=half_norms,
                                                    want_quantized=True)
        codes = codes.view(s, b, t)
        losses = torch.zeros(s, dtype=x.dtype, device=x.device)
    return quantized, codes, losses.view(s, 1)


class ResidualVectorQuantization(nn.Module):
    """Residual stack (reference core_vq.py:318-370): one fused kernel launch per call."""

    def __init__(self, *, num_quantizers, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([VectorQuantization(**kwargs) for _ in range(num_quantizers)])

    def _norms(self, st: int, n_q: int) -> torch.Tensor:
        """Norms are computed for the whole stack once and sliced per call."""
        return _half_norms(self, list(self.layers))[st:n_q]

    def _pack(self, st: int) -> tp.Optional[torch.Tensor]:
        """Tensor-core pack of the whole stack, offset to start at layer `st`."""
        pack = _tc_pack(self, list(self.layers))
        if pack is None or st == 0:
            return pack
        k, d = self.layers[0]._codebook.embed.shape
        return pack[st * ops.tc_pack_table_bytes(k, d):]

    def _fusable(self, layers) -> bool:
        return all(l._codebook.is_inited and not l.has_projection for l in layers)

    def forward(self, x: torch.Tensor, n_q: tp.Optional[int] = None):
        n_q = n_q or len(self.layers)
        layers = list(self.layers[:n_q])
        if self._fusable(layers):
            return _stack_forward(layers, x, self.training, self._norms(0, n_q), self._pack(0),
                                  all_layers=list(self.layers))
        # one-off path: a codebook still needs its k-means initialisation (first forward of a
        # kmeans_init=True module, core_vq.py:207) or carries a projection: go layer by layer
        quantized_out = 0.0
        residual = x
        all_losses, all_indices = [], []
        for layer in layers:
            quantized, indices, loss = layer(residual)
            residual = residual - quantized
            quantized_out = quantized_out + quantized
            all_indices.append(indices)
            all_losses.append(loss)
        return quantized_out, torch.stack(all_indices), torch.stack(all_losses)

    def encode(self, x: torch.Tensor, n_q: tp.Optional[int] = None,
               st: tp.Optional[int] = None) -> torch.Tensor:
        n_q = n_q or len(self.layers)
        st = st or 0
        layers = list(self.layers[st:n_q])
        b, _, t = x.shape
        if not any(l.has_projection for l in layers):
            codes, _, _, _ = ops.rvq_search(x, [l._codebook.embed for l in layers], len(layers),
                                            half_norms=self._norms(st, n_q), tc_pack=self._pack(st))
            return codes.view(len(layers), b, t)
        residual, out = x, []
        for layer in layers:
            idx = layer.encode(residual)
            residual = residual - layer.decode(idx)
            out.append(idx)
        return torch.stack(out)

    def decode(self, q_indices: torch.Tensor) -> torch.Tensor:
        s, b, t = q_indices.shape
        layers = list(self.layers[:s])
        if not any(l.has_projection for l in layers):
            return ops.vq_decode(q_indices, b * t, 1, [l._codebook.embed for l in layers], s, 1, b, t)
        out = torch.tensor(0.0, device=q_indices.device)
        for i, idx in enumerate(q_indices):
            out = out + self.layers[i].decode(idx)
        return out
