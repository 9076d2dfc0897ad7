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
