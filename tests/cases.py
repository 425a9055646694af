"""Seeded parity cases shared by tests/golden/make_golden.py (which runs the live reference)
and the tests (which run the oracle and the CUDA path).  Inputs are rebuilt from seeds by
academicodec_b200.synth, so the fixtures only hold the reference's *outputs*."""
from __future__ import annotations

import numpy as np
import torch

from academicodec_b200 import synth

RVQ_CASES = {
    # name: (D, n_q, bins, B, T, frame_rate, regime, x_seed, cb_seed)
    "cfg1_small": dict(D=128, n_q=8, bins=1024, B=2, T=100, frame_rate=100, regime="decay",
                       x_seed=1234, cb_seed=4321),
    "cfg1_randn": dict(D=128, n_q=8, bins=1024, B=3, T=77, frame_rate=100, regime="randn",
                       x_seed=99, cb_seed=5000),
    "recipe_d512": dict(D=512, n_q=12, bins=1024, B=1, T=37, frame_rate=100, regime="decay",
                        x_seed=7, cb_seed=8000),
    "vq1_750fps": dict(D=512, n_q=2, bins=1024, B=1, T=101, frame_rate=750, regime="randn",
                       x_seed=21, cb_seed=8100),
    "odd_dims": dict(D=40, n_q=3, bins=200, B=2, T=13, frame_rate=100, regime="decay",
                     x_seed=5, cb_seed=8200),
}

GRVQ_CASES = {
    "grvq_randn": dict(G=2, n_codes=1024, B=3, T=50, regime="randn", x_seed=31, cb_seed=777),
    "grvq_init": dict(G=2, n_codes=1024, B=2, T=19, regime="init", x_seed=32, cb_seed=778,
                      x_scale=0.01),
    "grvq_g4": dict(G=4, n_codes=256, B=2, T=33, regime="randn", x_seed=33, cb_seed=779),
}


def rvq_inputs(case: dict):
    x = torch.from_numpy(synth.latents(case["B"], case["D"], case["T"], case["x_seed"]))
    cb = torch.from_numpy(synth.rvq_codebooks(case["n_q"], case["bins"], case["D"],
                                              case["cb_seed"], case["regime"]))
    return x, cb


def tie_inputs():
    """Exact-tie stress: stage 0 all-zero codebook (every distance ties -> index 0, the state a
    kmeans_init=True module is in before its first training forward); stage 1 has
    duplicated rows (first duplicate must win); stage 2 ordinary."""
    d, k = 32, 64
    x = torch.from_numpy(synth.latents(2, d, 21, 404))
    cb = torch.from_numpy(synth.rvq_codebooks(3, k, d, 405, "randn")).clone()
    cb[0].zero_()
    cb[1, 32:] = cb[1, :32]
    return x, cb


def grvq_inputs(case: dict):
    x = torch.from_numpy(synth.latents(case["B"], 512, case["T"], case["x_seed"],
                                       case.get("x_scale", 1.0)))
    w = synth.grvq_codebooks(case["G"], case["n_codes"], case["cb_seed"], case["regime"])
    w = [[torch.from_numpy(g) for g in stage] for stage in w]
    return x, w


def kmeans_inputs():
    n, k, d, iters = 500, 64, 16, 5
    samples = torch.from_numpy(synth.normal((n, d), 606))
    return samples, k, iters


BITPACK_CASES = [(1, 1), (1, 9), (7, 8), (7, 13), (10, 1), (10, 8), (10, 1000), (10, 4097), (15, 33), (16, 100)]


def bitpack_values(bits: int, n: int) -> np.ndarray:
    """Seeded codes in [0, 2^bits)."""
    u = synth.uniform((n,), 9000 + 31 * bits + n, 0.0, 1.0).astype(np.float64)
    return np.minimum((u * (1 << bits)).astype(np.int64), (1 << bits) - 1)
