"""torch.distributed helpers used by the quantizer (mirrors the live parts of the reference's
quantization/distrib.py: `broadcast_tensors` :56-71 and `all_reduce` :30-32)."""
from __future__ import annotations

import os
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


class PeerExchange:
    """In-place all-reduce(SUM) of one flat fp32 buffer over NVLink peer memory (acq_peer_allreduce).

    The buffer lives in torch symmetric memory: every rank maps every rank's copy and -- on NVSwitch systems --
    one multicast address covering all of them, so the two-shot all-reduce is ONE kernel of this package
    (multimem.ld_reduce + multimem.st through the switch; plain peer loads / stores otherwise) between two
    cross-rank barriers of the symmetric-memory handle.  Construction is collective (all ranks of the group must
    create the exchange at the same point with the same size).  `ACQ_PEER_REDUCE=0`, or a platform without
    symmetric memory, leaves `PeerExchange.create` returning None and the caller on NCCL."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self.numel = (numel + 3) // 4 * 4
        self.buf = symm.empty(self.numel, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.world = int(self.handle.world_size)
        self.rank = int(self.handle.rank)
        mc = 0
        try:
            if os.environ.get("ACQ_PEER_MULTICAST", "1") != "0" and self.handle.has_multicast_support:
                mc = int(self.handle.multicast_ptr)
        except Exception:
            mc = 0
        self.multicast_ptr = mc
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.mode = "nvls-multimem" if mc else "p2p"

    @classmethod
    def create(cls, numel: int, device: torch.device, group=None) -> tp.Optional["PeerExchange"]:
        if not is_distributed() or os.environ.get("ACQ_PEER_REDUCE", "1") == "0":
            return None
        if dist.get_backend(group) != "nccl" or device.type != "cuda":
            return None
        try:
            return cls(numel, device, group)
        except Exception as exc:          # no symmetric memory on this platform: NCCL serves the collective
            import warnings
            warnings.warn(f"academicodec_b200: peer-memory exchange unavailable ({exc!r}); using NCCL all_reduce")
            return None

    def all_reduce_(self) -> torch.Tensor:
        """Sum self.buf over the ranks in place (stream-ordered on the current stream); returns self.buf."""
        import ctypes
        from .. import _lib
        lib = _lib.load()
        self.handle.barrier(channel=0)                         # every rank's statistics are written
        arr = (ctypes.c_void_p * self.world)(*self.peer_ptrs)
        stream = torch.cuda.current_stream(self.buf.device).cuda_stream
        with torch.cuda.device(self.buf.device):
            rc = lib.acq_peer_allreduce(ctypes.c_void_p(self.multicast_ptr) if self.multicast_ptr else None,
                                        ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p)), self.world, self.rank,
                                        self.numel, ctypes.c_void_p(stream))
        _lib.check(rc, "acq_peer_allreduce")
        self.handle.barrier(channel=1)                         # every rank's slice has been stored everywhere
        return self.buf
