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
