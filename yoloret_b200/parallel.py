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


class GradBucket:
    """Data-parallel gradient exchange for the training config (reference: MirroredStrategy's implicit
    all-reduce, code/train.py:55-56, and the loss SUM of code/yolo3/train.py:66-70).

    All gradients live in ONE flat fp32 bucket (2.63 M elements = 10.5 MB for MobileNetV2-0.75 COCO), padded
    to a multiple of the world size.  ``reduce_scatter`` leaves each rank with the SUM of its 1/N shard
    (``ncclReduceScatter`` over NVLink; a sharded optimizer updates that shard), ``all_gather`` rebuilds the
    full vector (updated parameters).  The reference sums across replicas (it does not average, each
    replica divides by its LOCAL batch, code/yolo3/model.py:624-625) - so does this."""

    def __init__(self, numel: int, world: int, rank: int, device=None, group=None):
        if numel <= 0 or world <= 0 or not (0 <= rank < world):
            raise ValueError("bad bucket: numel=%d world=%d rank=%d" % (numel, world, rank))
        self.numel, self.world, self.rank, self.group = numel, world, rank, group
        self.shard_numel = (numel + world - 1) // world
        self.padded = self.shard_numel * world
        self.flat = torch.zeros(self.padded, dtype=torch.float32, device=device)
        self.shard = torch.zeros(self.shard_numel, dtype=torch.float32, device=device)

    def view(self) -> torch.Tensor:
        """The caller-visible gradient vector (without the padding)."""
        return self.flat[:self.numel]

    def reduce_scatter(self) -> torch.Tensor:
        if self.world == 1:
            self.shard.copy_(self.flat[:self.shard_numel])
        elif self.flat.is_cuda:
            dist.reduce_scatter_tensor(self.shard, self.flat, op=dist.ReduceOp.SUM, group=self.group)
        else:  # gloo has no reduce_scatter: all-reduce then slice (CPU tests only)
            tmp = self.flat.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            self.shard.copy_(tmp[self.rank * self.shard_numel:(self.rank + 1) * self.shard_numel])
        return self.shard

    def all_gather(self) -> torch.Tensor:
        if self.world == 1:
            self.flat[:self.shard_numel].copy_(self.shard)
        elif self.flat.is_cuda:
            dist.all_gather_into_tensor(self.flat, self.shard, group=self.group)
        else:
            dist.all_gather(list(self.flat.split(self.shard_numel)), self.shard, group=self.group)
        return self.flat[:self.numel]
