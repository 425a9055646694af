"""Platform-independent synthetic inputs for parity tests, golden fixtures and the bench.

Every value is produced by integer arithmetic (a splitmix64 counter hash) followed by one
IEEE multiply, so the bytes are identical on any host, numpy or torch version.  The four
16-bit fields of each hash are summed (Irwin-Hall, n=4) which gives an approximately normal
variate with unit variance after scaling; the values sit on a 2^-16-ish grid and are exactly
representable in fp32.

Workload definitions follow SURVEY.md section 8(d) (cfg1..cfg5).
"""
from __future__ import annotations

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
# std of the sum of four independent uniform 16-bit integers
_IH4_STD = float(np.sqrt(4.0 * ((2.0**32 - 1.0) / 12.0)))


def _splitmix64(counter: np.ndarray) -> np.ndarray:
    z = counter
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def normal(shape, seed: int, scale: float = 1.0, chunk: int = 1 << 24) -> np.ndarray:
    """Approximately N(0, scale^2) fp32 array of `shape`, deterministic in (seed, index)."""
    n = int(np.prod(shape))
    out = np.empty(n, dtype=np.float32)
    base = np.uint64((int(seed) * int(_GOLDEN)) & 0xFFFFFFFFFFFFFFFF)
    mul = np.float64(scale) / np.float64(_IH4_STD)
    with np.errstate(over="ignore"):
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            ctr = (np.arange(lo, hi, dtype=np.uint64) + np.uint64(1)) * _GOLDEN + base
            h = _splitmix64(ctr)
            s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
                 + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48)))
            out[lo:hi] = ((s.astype(np.float64) - 131070.0) * mul).astype(np.float32)
    return out.reshape(shape)


def uniform(shape, seed: int, lo: float, hi: float) -> np.ndarray:
    """Uniform fp32 in [lo, hi) from the top 24 bits of the hash."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        ctr = (np.arange(n, dtype=np.uint64) + np.uint64(1)) * _GOLDEN \
            + np.uint64((int(seed) * int(_GOLDEN)) & 0xFFFFFFFFFFFFFFFF)
        h = _splitmix64(ctr) >> np.uint64(40)
    u = h.astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def latents(batch: int, dim: int, frames: int, seed: int = 1234, scale: float = 1.0) -> np.ndarray:
    """Encoder latents `[B, D, T]` fp32 (the layout the reference models hand to the quantizer,
    reference net3.py:39-43)."""
    return normal((batch, dim, frames), seed, scale)


def rvq_codebooks(n_q: int, bins: int, dim: int, seed: int = 4321, regime: str = "randn",
                  shrink: float = 0.7) -> np.ndarray:
    """Codebooks `[n_q, bins, dim]` fp32.

    regime "randn":  every stage ~ N(0,1)  (stresses near-ties uniformly, SURVEY 8d regime i)
    regime "decay":  stage s ~ N(0, shrink^(2s)) so later stages match shrinking residuals
    """
    cb = np.empty((n_q, bins, dim), dtype=np.float32)
    for s in range(n_q):
        sc = 1.0 if regime == "randn" else shrink ** s
        cb[s] = normal((bins, dim), seed + s, sc)
    return cb


def grvq_codebooks(n_groups: int, n_codes: int, seed: int = 777, regime: str = "randn"):
    """Two residual stages x `n_groups` codebooks `[n_codes, 512 // n_groups]`
    (reference hificodec/models.py:446-456).  regime "init" reproduces the reference's
    U(-1/n_codes, 1/n_codes) initialisation range (models.py:433-434)."""
    dg = 512 // n_groups
    out = []
    for stage in range(2):
        row = []
        for g in range(n_groups):
            sd = seed + 10 * stage + g
            if regime == "init":
                row.append(uniform((n_codes, dg), sd, -1.0 / n_codes, 1.0 / n_codes))
            else:
                row.append(normal((n_codes, dg), sd, 1.0 if stage == 0 else 0.6))
        out.append(row)
    return out


# ---- named workloads ---------------------------------------------------------
WORKLOADS = {
    # name: dict(kind, D, n_q, bins, frame_rate, B, T)
    "cfg1_enc24k_240d_rvq": dict(kind="rvq", D=128, n_q=8, bins=1024, frame_rate=100, B=16, T=100),
    "cfg1_recipe_rvq": dict(kind="rvq", D=512, n_q=12, bins=1024, frame_rate=100, B=16, T=100),
    "cfg2_enc24k_32d_vq1": dict(kind="rvq", D=512, n_q=1, bins=1024, frame_rate=750, B=8, T=45000),
    "cfg3_hifi16k_320d_grvq": dict(kind="grvq", D=512, G=2, n_q=2, bins=1024, frame_rate=50, B=64, T=50),
    "cfg4_ss24k_240d_rvq": dict(kind="rvq", D=512, n_q=12, bins=1024, frame_rate=100, B=8, T=1000),
    "cfg5_rvq_ema_train": dict(kind="rvq_train", D=128, n_q=8, bins=1024, frame_rate=100, B=16, T=100),
    # throughput variants of the same shapes (B in {16, 256, 4096} per GPU, 64 x 10 s, 4096 x 50)
    "cfg1_b4096": dict(kind="rvq", D=128, n_q=8, bins=1024, frame_rate=100, B=4096, T=100),
    "cfg3_b4096": dict(kind="grvq", D=512, G=2, n_q=2, bins=1024, frame_rate=50, B=4096, T=50),
    "cfg4_b64": dict(kind="rvq", D=512, n_q=12, bins=1024, frame_rate=100, B=64, T=1000),
    "cfg5_recipe_train": dict(kind="rvq_train", D=512, n_q=12, bins=1024, frame_rate=100, B=16, T=100),
    "cfg5_b640_train": dict(kind="rvq_train", D=512, n_q=12, bins=1024, frame_rate=100, B=640, T=100),
}
