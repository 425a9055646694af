"""B200-native group-residual VQ: drop-in for the reference HiFi-Codec `Quantizer` /
`Quantizer_module` (academicodec/models/hificodec/models.py:430-535).

Same constructor (`Quantizer(h)` with h.n_code_groups, h.n_codes, h.codebook_loss_lambda,
h.commitment_loss_lambda), same parameters / state-dict keys
(`quantizer_modules.{g}.embedding.weight`, `quantizer_modules2.{g}.embedding.weight`), same
`forward(xin) -> (quantized_out, loss, [indices...])` and `embed(codes) -> [B, 512, T]`.
Both residual stages and all groups run in ONE launch of the fused search kernel."""
from __future__ import annotations

import torch
from torch import nn

from . import ops

CHANNELS = 512   # hard-coded in the reference (models.py:448,450,465-466)
_EAGER_BACKWARD = False   # tests: take the eager expression of the backward pass instead of acq_grvq_backward


class Quantizer_module(nn.Module):
    """One codebook as an nn.Embedding parameter, U(-1/n_e, 1/n_e) init (models.py:430-434)."""

    def __init__(self, n_e: int, e_dim: int):
        super().__init__()
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def forward(self, x: torch.Tensor):
        """x [N, e_dim] -> (z_q [N, e_dim], indices [N]) (models.py:436-442)."""
        w = self.embedding.weight.detach()
        codes, _, _, _ = ops.rvq_search(x.detach().t().contiguous().unsqueeze(0), [w], 1)
        idx = codes.view(-1)
        return self.embedding(idx), idx


class _GroupResidualSearch(torch.autograd.Function):
    """Forward: fused 2-stage x G-group search with the reference's always-on straight-through
    arithmetic.  Backward: the reference's gradients -- identity to xin from the quantized sum,
    commitment term to xin from stage 0 only, codebook term scattered into every codebook -- one kernel
    (acq_grvq_backward); the eager expression of the same gradients below it serves shapes the kernel does not
    take (channel counts that are not a multiple of 32)."""

    @staticmethod
    def forward(ctx, xin, lam_cb, lam_commit, n_groups, pack, *weights):
        stages = len(weights) // n_groups
        ws = [w.detach() for w in weights]
        b, c, t = xin.shape
        flags = ops.ACQ_STE | ops.ACQ_LOSS_RAW
        if pack is not None:
            # tensor-core search for the codes + one replay pass for quantized / loss
            codes, _, _, _ = ops.rvq_search(xin.detach(), ws, stages, n_groups, flags=flags, tc_pack=pack)
            quantized, _, sqerr, _ = ops.rvq_replay(xin.detach(), codes, ws, stages, n_groups, flags=flags,
                                                    want_sqerr=True)
        else:
            codes, quantized, _, sqerr = ops.rvq_search(
                xin.detach(), ws, stages, n_groups, flags=flags, want_quantized=True, want_sqerr=True)
        # loss_s = lam_cb * mean((zq - x)^2) + lam_commit * mean((zq - x)^2)   (models.py:476-477)
        losses = (sqerr * ((lam_cb + lam_commit) / float(xin.numel()))).to(xin.dtype)
        ctx.save_for_backward(xin, codes, *weights)
        ctx.cfg = (lam_cb, lam_commit, n_groups, stages)
        ctx.mark_non_differentiable(codes)
        return quantized, losses, codes

    @staticmethod
    def backward(ctx, g_q, g_losses, _g_codes):
        xin, codes, *weights = ctx.saved_tensors
        lam_cb, lam_commit, n_groups, stages = ctx.cfg
        b, c, t = xin.shape
        if xin.is_cuda and c % 32 == 0 and c <= 768 and (c // n_groups) % 32 == 0 and not _EAGER_BACKWARD:
            # one kernel: residual chain recomputed from x and the codes, d xin and the scattered codebook
            # gradients (acq_grvq_backward)
            grad_x, grad_w = ops.grvq_backward(
                xin.detach(), codes, [w.detach() for w in weights], stages, n_groups, g_q, g_losses, lam_cb, lam_commit,
                want_grad_x=ctx.needs_input_grad[0],
                want_grad_cb=[bool(ctx.needs_input_grad[5 + i]) and g_losses is not None for i in range(len(weights))])
            return (grad_x, None, None, None, None, *grad_w)
        dg = c // n_groups
        numel = float(xin.numel())
        r = xin.detach().transpose(1, 2).reshape(-1, c)           # [N, 512] residual entering stage s
        grad_x = g_q.clone() if g_q is not None else torch.zeros_like(xin)
        grad_w = [None] * len(weights)
        for s in range(stages):
            zq = torch.cat([weights[s * n_groups + g].detach()[codes[s * n_groups + g]]
                            for g in range(n_groups)], -1)          # [N, 512]
            diff = zq - r
            if g_losses is not None:
                gl = g_losses[s]
                if s == 0 and ctx.needs_input_grad[0]:
                    # d/dx of lam_commit * mean((zq.detach() - x)^2)
                    gx = (-2.0 * lam_commit / numel) * gl * diff
                    grad_x = grad_x + gx.view(b, t, c).transpose(1, 2)
                for g in range(n_groups):
                    i = s * n_groups + g
                    if ctx.needs_input_grad[5 + i]:
                        gw = torch.zeros_like(weights[i])
                        gw.index_add_(0, codes[i], (2.0 * lam_cb / numel) * gl
                                      * diff[:, g * dg:(g + 1) * dg])
                        grad_w[i] = gw
            r = r - (r + (zq - r))                                   # next stage's residual
        return (grad_x, None, None, None, None, *grad_w)


class Quantizer(nn.Module):
    def __init__(self, h):
        super().__init__()
        assert CHANNELS % h.n_code_groups == 0
        self.quantizer_modules = nn.ModuleList([
            Quantizer_module(h.n_codes, CHANNELS // h.n_code_groups) for _ in range(h.n_code_groups)])
        self.quantizer_modules2 = nn.ModuleList([
            Quantizer_module(h.n_codes, CHANNELS // h.n_code_groups) for _ in range(h.n_code_groups)])
        self.h = h
        self.codebook_loss_lambda = self.h.codebook_loss_lambda
        self.commitment_loss_lambda = self.h.commitment_loss_lambda
        self.residul_layer = 2      # (sic) attribute name kept from the reference, models.py:460
        self.n_code_groups = h.n_code_groups

    def _weights(self):
        return [m.embedding.weight for m in self.quantizer_modules] + \
               [m.embedding.weight for m in self.quantizer_modules2]

    def forward(self, xin: torch.Tensor):
        """xin [B, 512, T] -> (quantized_out [B, 512, T], loss 0-d,
        [idx_s0g0, .., idx_s1g0, ..] each [B*T] int64)  (models.py:494-508)."""
        if xin.shape[1] != CHANNELS:
            raise RuntimeError(f"Quantizer expects {CHANNELS} channels, got {xin.shape[1]}")
        quantized, losses, codes = _GroupResidualSearch.apply(
            xin, float(self.codebook_loss_lambda), float(self.commitment_loss_lambda),
            self.n_code_groups, self._pack(xin), *self._weights())
        return quantized, torch.mean(losses), list(codes.unbind(0))

    def _pack(self, xin: torch.Tensor):
        """Tensor-core operand pack of the four codebooks, rebuilt when a weight changes (optimizer
        steps bump the tensors' version counters); None when the kernel does not apply."""
        ws = self._weights()
        k = ws[0].shape[0]
        if not (xin.is_cuda and ops.tc_supported(k, CHANNELS, self.n_code_groups)):
            return None
        key = tuple((w.data_ptr(), w._version, w.device) for w in ws) + (getattr(self, "_cache_epoch", 0),)
        cached = getattr(self, "_tc_cache", None)
        if cached is None or cached[0] != key:
            cached = (key, ops.tc_pack_codebooks([w.detach() for w in ws]))
            self._tc_cache = cached
        return cached[1]

    def invalidate_caches(self) -> None:
        """Call after writing a codebook through `.data` (e.g. `embedding.weight.data.copy_(w)`): such
        writes do not bump the parameter's version counter, which is what the operand-pack cache keys on."""
        self._cache_epoch = getattr(self, "_cache_epoch", 0) + 1

    @torch.no_grad()
    def encode(self, xin: torch.Tensor):
        """Codes only: the list `forward` returns as its third element, without quantized / loss.
        (Not in the reference; `VQVAE.encode`, vqvae.py:37-45, runs forward and discards the
        rest.)  Takes the tcgen05 kernel when the shape allows."""
        if xin.shape[1] != CHANNELS:
            raise RuntimeError(f"Quantizer expects {CHANNELS} channels, got {xin.shape[1]}")
        ws = [w.detach() for w in self._weights()]
        pack = self._pack(xin)
        codes, _, _, _ = ops.rvq_search(xin.detach(), ws, self.residul_layer, self.n_code_groups,
                                        flags=ops.ACQ_STE, tc_pack=pack)
        return list(codes.unbind(0))

    def embed(self, x: torch.Tensor) -> torch.Tensor:
        """codes [B, T, 2G] int64 (order s0g0, s0g1, .., s1g0, ..) -> [B, 512, T]
        (models.py:510-535)."""
        b, t, n = x.shape
        g = self.n_code_groups
        if n != self.residul_layer * g:
            raise RuntimeError(f"expected {self.residul_layer * g} code columns, got {n}")
        ws = [w.detach() for w in self._weights()]
        return ops.vq_decode(x, 1, n, ws, self.residul_layer, g, b, t)
