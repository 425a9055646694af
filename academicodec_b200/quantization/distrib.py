"""torch.distributed helpers used by the quantizer (mirrors the live parts of the reference's
quantization/distrib.py: `broadcast_tensors` :56-71 and `all_reduce` :30-32)."""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def is_distributed() -> bool:
    return world_size() > 1


def all_reduce(tensor: torch.Tensor, op=dist.ReduceOp.SUM if dist.is_available() else None):
    if is_distributed():
        return dist.all_reduce(tensor, op)


def broadcast_tensors(tensors: tp.Iterable[torch.Tensor], src: int = 0) -> None:
    """Make floating-point tensors identical on every rank (used after k-means init)."""
    if not is_distributed():
        return
    tensors = [t for t in tensors if torch.is_floating_point(t) or torch.is_complex(t)]
    count = torch.tensor([len(tensors)], device=tensors[0].device if tensors else "cpu", dtype=torch.long)
    dist.all_reduce(count)
    if count.item() != len(tensors) * world_size():
        raise RuntimeError(f"Mismatch in number of params: ours is {len(tensors)}, "
                           "at least one worker has a different one.")
    handles = [dist.broadcast(t.data, src=src, async_op=True) for t in tensors]
    for h in handles:
        h.wait()
