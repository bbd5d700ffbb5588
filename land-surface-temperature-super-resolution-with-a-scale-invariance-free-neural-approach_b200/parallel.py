"""Host-side data-parallel helpers (no arithmetic on the hot path): how patches are split
across ranks and how the flat gradient buffer is exchanged.

* Inference (reference predict.py:84-103): the patches of a tile are independent, so the
  patch list is block-partitioned across ranks with no collective at all
  (``block_partition(324, 8)`` -> 41,41,41,41,40,40,40,40).
* Training: one all-reduce of the 282 705-float gradient buffer per step, issued as two
  buckets (decoder half first, while the encoder backward is still running).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def block_partition(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous (start, count) per rank; the first ``n_items % world_size`` ranks get one extra item."""
    if world_size <= 0 or n_items < 0:
        raise ValueError("block_partition needs world_size > 0 and n_items >= 0")
    base, extra = divmod(n_items, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """This rank's contiguous slice of a global batch (dim 0)."""
    start, cnt = block_partition(t.shape[0], world_size)[rank]
    return t[start:start + cnt]


class BucketedAllReduce:
    """Two-bucket all-reduce of a flat gradient buffer.

    ``start(bucket)`` launches the asynchronous all-reduce of that bucket (0 = tail of the buffer, i.e. the
    decoder parameters whose gradients are final first; 1 = head / encoder); ``finish()`` waits for both.
    Works on CPU tensors with gloo (tests) and CUDA tensors with NCCL (product)."""

    def __init__(self, flat: torch.Tensor, split: int, group: Optional[dist.ProcessGroup] = None):
        if not 0 <= split <= flat.numel():
            raise ValueError("split outside the buffer")
        self.flat, self.split, self.group = flat, split, group
        self._work = []

    def start(self, bucket: int) -> None:
        view = self.flat[self.split:] if bucket == 0 else self.flat[:self.split]
        if view.numel():
            self._work.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        for w in self._work:
            w.wait()
        self._work = []
