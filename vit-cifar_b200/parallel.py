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


def exchange_peer_pointers(tensors, group=None):
    """Map every rank's copies of `tensors` (a dict name -> CUDA tensor, same shapes on all ranks of one node; None values are
    passed through) into this process through CUDA IPC.  Returns (dict name -> ops.PeerPointers in rank order, with this rank's
    own address at [rank]; error or None).  Exactly ONE collective is issued (an all_gather_object of the 64-byte handles), and
    it is issued before anything that can fail locally, so a rank whose mapping fails cannot desynchronise the others."""
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine, err = {}, None
    try:
        mine = {k: ops.ipc_export(t) for k, t in tensors.items() if t is not None}
    except Exception as e:  # e.g. memory that cannot be exported (expandable segments)
        err = e
    everyone = [None] * world
    dist.all_gather_object(everyone, mine if err is None else None, group=group)
    if err is None and any(e is None for e in everyone):
        err = RuntimeError("another rank could not export its buffers")
    out = {}
    if err is None:
        try:
            for k, t in tensors.items():
                if t is None:
                    out[k] = None
                    continue
                out[k] = ops.PeerPointers([t.data_ptr() if r == rank else ops.ipc_open(*everyone[r][k]) for r in range(world)])
        except Exception as e:  # no peer access between two of the GPUs
            err = e
    return (out if err is None else None), err
