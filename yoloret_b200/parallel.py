"""Multi-GPU inference: batch sharding + the all-gather of final detections.

The reference has no inference parallelism at all (single device, single image:
reference code/yolo.py:82-86,119) and trains data-parallel through
``tf.distribute.MirroredStrategy`` (code/train.py:55-56).  Images are independent, so
the batch is partitioned contiguously over ranks, weights are replicated, and the only
exchange is ONE NCCL all-gather per step of each rank's packed detection wire
(``PostProcess.wire``: counts | status | boxes | scores | classes), SURVEY.md §8e.
One process per GPU; ``torch.distributed`` is the plumbing.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a global batch owned by ``rank`` (remainder to the low ranks)."""
    if not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard request: batch=%d world=%d rank=%d" % (batch, world, rank))
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DetectionGather:
    """All-gathers equal-sized per-rank detection wires.  ``wire``: the rank's flat int32
    tensor (``PostProcess.wire``); on CUDA the collective is NCCL over NVLink, on CPU (tests)
    gloo."""

    def __init__(self, pp_or_wire, world: int, rank: int, group=None):
        self.wire = pp_or_wire.wire if hasattr(pp_or_wire, "wire") else pp_or_wire
        if self.wire.dtype != torch.int32 or self.wire.dim() != 1:
            raise ValueError("wire must be a flat int32 tensor")
        self.world, self.rank, self.group = world, rank, group
        self.words = self.wire.numel()
        self.gathered = torch.zeros(world * self.words, dtype=torch.int32, device=self.wire.device)
        self.host = torch.zeros(world * self.words, dtype=torch.int32)
        if self.wire.is_cuda:
            self.host = self.host.pin_memory()

    def all_gather(self) -> torch.Tensor:
        if self.wire.is_cuda:
            dist.all_gather_into_tensor(self.gathered, self.wire, group=self.group)
        else:
            dist.all_gather(list(self.gathered.split(self.words)), self.wire, group=self.group)
        return self.gathered

    def read(self) -> List[np.ndarray]:
        """Device->host copy of the gathered wires; one numpy view per rank, rank order = batch order."""
        self.host.copy_(self.gathered, non_blocking=True)
        if self.wire.is_cuda:
            torch.cuda.current_stream(self.wire.device).synchronize()
        return list(self.host.numpy().reshape(self.world, self.words))
