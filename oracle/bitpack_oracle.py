"""Oracle: code wire format.  TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's BitPacker / BitUnpacker (academicodec/binary.py:54-123):
values are appended LSB-first to one little-endian bit stream; the last partial byte is
flushed as is (binary.py:82-88)."""
from __future__ import annotations

import numpy as np


def pack(values, bits: int) -> bytes:
    v = np.asarray(values, dtype=np.uint64).reshape(-1)
    shifts = np.arange(bits, dtype=np.uint64)
    stream = ((v[:, None] >> shifts[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
    return np.packbits(stream, bitorder="little").tobytes()


def unpack(data: bytes, n: int, bits: int) -> np.ndarray:
    stream = np.unpackbits(np.frombuffer(data, dtype=np.uint8), bitorder="little")[: n * bits]
    weights = (np.uint64(1) << np.arange(bits, dtype=np.uint64))
    return (stream.reshape(n, bits).astype(np.uint64) * weights[None, :]).sum(1).astype(np.int64)
