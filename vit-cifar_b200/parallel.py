"""Data-parallel gradient exchange (what Lightning's DDP does for the reference when devices > 1, main.py:220-231;
SURVEY.md §8e): sum all-reduce of contiguous buckets of the flat gradient buffer, optionally on a side stream so it
overlaps the rest of backward.  The mean (1/world_size) is applied by the Adam kernel's gradient scale."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def allreduce_bucket(flat_grad: torch.Tensor, bucket: Tuple[int, int], group=None, comm_stream: Optional["torch.cuda.Stream"] = None) -> None:
    a, b = bucket
    if b <= a:
        return
    view = flat_grad[a:b]
    if comm_stream is not None:
        comm_stream.wait_stream(torch.cuda.current_stream())  # gradients of this bucket are complete on the compute stream
        with torch.cuda.stream(comm_stream):
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)


def allreduce_all(flat_grad: torch.Tensor, buckets: Sequence[Tuple[int, int]], group=None) -> None:
    """Backward order: head bucket first, stem last (the order gradients become ready)."""
    for bk in reversed(list(buckets)):
        allreduce_bucket(flat_grad, bk, group, None)


def owned_slice(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Element range of the flat buffers whose reduction and Adam update rank `rank` performs in the fused data-parallel step
    (vitb_dp_reduce_adam): ceil(n/4 / world) float4 groups per rank, in rank order."""
    n4 = n // 4
    per = (n4 + world - 1) // world
    return min(per * rank, n4) * 4, min(per * (rank + 1), n4) * 4


def exchange_peer_pointers(t: torch.Tensor, group=None):
    """Map every rank's copy of `t` (same shape on all ranks of one node) into this process through CUDA IPC; returns the device
    addresses in rank order (this rank's own address at [rank]).  torch.distributed only carries the 64-byte handles."""
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = ops.ipc_export(t)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    ptrs = []
    for r, (handle, offset) in enumerate(everyone):
        ptrs.append(t.data_ptr() if r == rank else ops.ipc_open(handle, offset))
    return ops.PeerPointers(ptrs)
