"""One-process-per-GPU data parallelism for the VCA-GAN step.

The reference's only multi-GPU mode is single-process nn.DataParallel (train.py:112-119): per-replica BatchNorm
statistics, gradients summed onto GPU 0, parameters re-broadcast on every forward.  Here every rank owns a full replica,
keeps its BatchNorm per replica (same semantics), and after each backward the flat gradient buffer of the D or G group
is sum-all-reduced in fixed-size buckets (NCCL over NVLink on GPUs, gloo in the CPU tests); the 1/world mean is folded
into the fused Adam kernel.  Nothing else crosses ranks.  This module is device agnostic so the host logic is testable
with world_size-2 gloo on CPU."""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous equal shards of the global batch (drop_last semantics of train.py:145: the batch must divide)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def bucket_ranges(numel: int, bucket_elems: int) -> List[Tuple[int, int]]:
    """[start, end) element ranges covering a flat buffer; the last bucket takes the ragged tail."""
    if bucket_elems <= 0:
        raise ValueError("bucket_elems must be positive")
    return [(s, min(s + bucket_elems, numel)) for s in range(0, numel, bucket_elems)]


def allreduce_flat(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None, bucket_elems: int = 8 << 20,
                   async_op: bool = False):
    """Sum all-reduce of a flat buffer in buckets (in place).  Returns the list of work handles when async_op."""
    works = []
    for s, e in bucket_ranges(flat.numel(), bucket_elems):
        w = dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def broadcast_flat(flat: torch.Tensor, src: int = 0, group: Optional[dist.ProcessGroup] = None):
    """Make every replica start from rank `src`'s weights (done once; replicas then stay identical because every rank
    applies the same averaged gradient)."""
    dist.broadcast(flat, src, group=group)


class NativeComm:
    """The library's own NCCL communicator (include/vcagan.h: vca_comm_unique_id / vca_comm_init / vca_allreduce_bucket), one
    per process and GPU.  torch.distributed is used ONCE, as the side channel that ships rank 0's 128-byte rendezvous token;
    the gradient exchange itself then runs through the C ABI (what a non-PyTorch host would call)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        import ctypes
        from ._lib import lib, VcaError
        self._lib = lib()
        if self._lib.cdll.vca_comm_available() != 1:
            raise VcaError("NativeComm: no NCCL library could be bound in this process")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        tok = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0 and self._lib.cdll.vca_comm_unique_id(ctypes.c_void_p(tok.data_ptr())) != 0:
            raise VcaError(self._lib.cdll.vca_last_error().decode())
        tok = tok.to(dev)
        dist.broadcast(tok, 0, group=group)
        tok = tok.cpu()
        with torch.cuda.device(dev):
            if self._lib.cdll.vca_comm_init(ctypes.c_void_p(tok.data_ptr()), self.rank, self.world) != 0:
                raise VcaError(self._lib.cdll.vca_last_error().decode())

    def allreduce_flat(self, flat: torch.Tensor, bucket_elems: int = 8 << 20):
        """In-place sum all-reduce of a flat fp32 / bf16 CUDA buffer in buckets, on the current stream."""
        from .ops import BF16, F32
        dt = {torch.float32: F32, torch.bfloat16: BF16}[flat.dtype]
        for s, e in bucket_ranges(flat.numel(), bucket_elems):
            self._lib.call("vca_allreduce_bucket", flat[s:e], e - s, dt)

    def close(self):
        self._lib.cdll.vca_comm_destroy()
