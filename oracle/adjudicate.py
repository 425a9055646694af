"""fp64 arbiter for codeword choices.  TEST INFRASTRUCTURE ONLY.

The reference's nearest-codeword search runs in fp32 (core_vq.py:175-180,
hificodec/models.py:436-441) and its sgemm summation order differs between hosts and
devices, so two correct implementations may legitimately disagree on a frame whose two best
distances are closer than fp32 can resolve.  This module decides, in float64, whether a
disagreement is such a near-tie or a real error (SURVEY.md section 8c "parity rules").

Stated tolerance:  a choice i is accepted for residual r iff
    d64(r, e_i) - min_k d64(r, e_k)  <=  EPS_ULPS * 2^-23 * (||r||^2 + max_k ||e_k||^2)
i.e. EPS_ULPS fp32 ulps of the magnitude the reference's distance expression is evaluated at
(||x||^2 - 2 x.e + ||e||^2 is rounded at that magnitude).  EPS_ULPS = 8.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

EPS_ULPS = 8.0
_ULP32 = 2.0 ** -23


def _f64(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def _f32(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float32)


def distances64(flat: np.ndarray, embed: np.ndarray) -> np.ndarray:
    """Exact-ish squared distances [N,K] in float64."""
    r = _f64(flat)
    e = _f64(embed)
    return (r * r).sum(1, keepdims=True) - 2.0 * (r @ e.T) + (e * e).sum(1)[None, :]


def tolerance(flat: np.ndarray, embed: np.ndarray) -> np.ndarray:
    r = _f64(flat)
    e = _f64(embed)
    return EPS_ULPS * _ULP32 * ((r * r).sum(1) + (e * e).sum(1).max())


def judge_choice(flat, embed, chosen, chunk: int = 8192) -> Dict[str, np.ndarray]:
    """Per-frame verdict of `chosen` [N] against the fp64 optimum.
    Returns excess (d_chosen - d_min), tol, best (fp64 argmin, first index)."""
    flat = _f32(flat)
    chosen = np.asarray(chosen).astype(np.int64).reshape(-1)
    n = flat.shape[0]
    excess = np.empty(n)
    tol = np.empty(n)
    best = np.empty(n, dtype=np.int64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        d = distances64(flat[lo:hi], embed)
        best[lo:hi] = d.argmin(1)
        dmin = d.min(1)
        excess[lo:hi] = d[np.arange(hi - lo), chosen[lo:hi]] - dmin
        tol[lo:hi] = tolerance(flat[lo:hi], embed)
    return dict(excess=excess, tol=tol, best=best)


def audit_rvq_codes(x_bdt, embeds: Sequence, codes_sbt, st: int = 0,
                    straight_through: bool = False) -> Dict[str, object]:
    """Walk an RVQ code sequence and verify every stage's choice is fp64-optimal within the
    stated tolerance *for the residual its own earlier choices produce* (the residual is
    propagated in fp32 exactly as the reference does: r - e[i], core_vq.py:359; or
    r - (r + (e[i] - r)) under the straight-through form, core_vq.py:304,339).

    -> dict(exact=[per stage], near_tie=[...], wrong=[...], worst_excess_over_tol=float)"""
    x = _f32(x_bdt)
    b, d, t = x.shape
    r = np.ascontiguousarray(x.transpose(0, 2, 1)).reshape(-1, d)
    codes = np.asarray(codes_sbt.cpu() if hasattr(codes_sbt, "cpu") else codes_sbt).astype(np.int64)
    exact, near, wrong = [], [], []
    worst = 0.0
    for j in range(codes.shape[0]):
        e = _f32(embeds[st + j])
        c = codes[j].reshape(-1)
        v = judge_choice(r, e, c)
        is_exact = v["excess"] <= 0.0
        is_near = (~is_exact) & (v["excess"] <= v["tol"])
        is_wrong = v["excess"] > v["tol"]
        exact.append(int(is_exact.sum()))
        near.append(int(is_near.sum()))
        wrong.append(int(is_wrong.sum()))
        worst = max(worst, float((v["excess"] / v["tol"]).max()))
        q = e[c]
        if straight_through:
            q = (r + (q - r)).astype(np.float32)
        r = (r - q).astype(np.float32)
    return dict(exact=exact, near_tie=near, wrong=wrong, worst_excess_over_tol=worst)


def compare_rvq_codes(x_bdt, embeds: Sequence, codes_ref, codes_new, st: int = 0,
                      straight_through: bool = False) -> Dict[str, object]:
    """Index-for-index comparison with near-tie classification.

    For every frame, find the first stage at which the two code sequences differ; re-evaluate
    both candidates in fp64 on the *reference's* residual at that stage.  The disagreement is
    a near-tie iff |d64[i_ref] - d64[i_new]| <= tolerance.  Later stages of such a frame
    follow different residuals and are reported as `downstream` (they are covered by
    audit_rvq_codes on the new sequence)."""
    x = _f32(x_bdt)
    b, d, t = x.shape
    r = np.ascontiguousarray(x.transpose(0, 2, 1)).reshape(-1, d)
    cr = np.asarray(codes_ref.cpu() if hasattr(codes_ref, "cpu") else codes_ref).astype(np.int64)
    cn = np.asarray(codes_new.cpu() if hasattr(codes_new, "cpu") else codes_new).astype(np.int64)
    assert cr.shape == cn.shape, (cr.shape, cn.shape)
    s = cr.shape[0]
    n = r.shape[0]
    diverged = np.zeros(n, dtype=bool)
    identical = near = hard = downstream = 0
    hard_examples: List[tuple] = []
    for j in range(s):
        e = _f32(embeds[st + j])
        a = cr[j].reshape(-1)
        c = cn[j].reshape(-1)
        diff = a != c
        downstream += int((diff & diverged).sum())
        fresh = diff & ~diverged
        identical += int((~diff).sum())
        if fresh.any():
            idx = np.nonzero(fresh)[0]
            rr = r[idx].astype(np.float64)
            ea = e[a[idx]].astype(np.float64)
            ec = e[c[idx]].astype(np.float64)
            da = ((rr - ea) ** 2).sum(1)
            dc = ((rr - ec) ** 2).sum(1)
            tol = tolerance(r[idx], e)
            ok = np.abs(da - dc) <= tol
            near += int(ok.sum())
            hard += int((~ok).sum())
            for k in np.nonzero(~ok)[0][:5]:
                hard_examples.append((j, int(idx[k]), int(a[idx[k]]), int(c[idx[k]]),
                                      float(da[k]), float(dc[k]), float(tol[k])))
            diverged |= fresh
        q = e[a]
        if straight_through:
            q = (r + (q - r)).astype(np.float32)
        r = (r - q).astype(np.float32)
    return dict(total=int(s * n), identical=identical, near_tie=near, hard_mismatch=hard,
                downstream=downstream, diverged_frames=int(diverged.sum()),
                hard_examples=hard_examples)
