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
    gloo.

    Two ways to drive it:
      * ``all_gather()`` [+ ``read()``]: the collective on the caller's stream, then a blocking device->host
        read - simple, serialises with the next step.
      * ``gather_async()`` + ``wait(ticket)``: the pipelined form.  The wire is snapshotted into a staging buffer
        (the only thing the compute stream waits for, ~2.5 MB device-to-device), then the all-gather and the
        device->host copy of the gathered buffer run on a side stream and overlap the next step's compute; the
        caller picks the result up one step later.  Device and pinned host buffers are double-buffered, so step
        i+1 never overwrites what step i's reader still holds."""

    def __init__(self, pp_or_wire, world: int, rank: int, group=None):
        self.wire = pp_or_wire.wire if hasattr(pp_or_wire, "wire") else pp_or_wire
        if self.wire.dtype != torch.int32 or self.wire.dim() != 1:
            raise ValueError("wire must be a flat int32 tensor")
        self.world, self.rank, self.group = world, rank, group
        self.words = self.wire.numel()
        dev = self.wire.device
        self._gathered = [torch.zeros(world * self.words, dtype=torch.int32, device=dev) for _ in range(2)]
        self._host = [torch.zeros(world * self.words, dtype=torch.int32) for _ in range(2)]
        self.cuda = self.wire.is_cuda
        if self.cuda:
            self._host = [h.pin_memory() for h in self._host]
            self._stage = torch.zeros_like(self.wire)
            self._side = torch.cuda.Stream(dev)
            self._staged = torch.cuda.Event()
            self._landed = [torch.cuda.Event(), torch.cuda.Event()]
        self._n = 0
        self.gathered = self._gathered[0]
        self.host = self._host[0]

    def _collective(self, dst: torch.Tensor, src: torch.Tensor):
        if self.cuda:
            dist.all_gather_into_tensor(dst, src, group=self.group)
        else:
            dist.all_gather(list(dst.split(self.words)), src, group=self.group)

    def all_gather(self) -> torch.Tensor:
        """The collective on the current stream; returns the gathered device tensor [world * words]."""
        self.gathered = self._gathered[self._n & 1]
        self._n += 1
        self._collective(self.gathered, self.wire)
        return self.gathered

    def read(self) -> List[np.ndarray]:
        """Blocking device->host copy of the last ``all_gather``; one numpy view per rank, rank order = batch order."""
        self.host.copy_(self.gathered, non_blocking=True)
        if self.cuda:
            torch.cuda.current_stream(self.wire.device).synchronize()
        return list(self.host.numpy().reshape(self.world, self.words))

    def gather_async(self, read: bool = True) -> int:
        """Pipelined form: snapshot the wire, then all-gather (+ device->host copy when ``read``) on a side stream.
        Returns a ticket for ``wait``.  The caller's stream only waits for the snapshot."""
        k = self._n & 1
        self._n += 1
        self.gathered = self._gathered[k]
        if not self.cuda:
            self._collective(self.gathered, self.wire)
            if read:
                self._host[k].copy_(self.gathered)
            return k
        main = torch.cuda.current_stream(self.wire.device)
        side = self._side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._stage.copy_(self.wire, non_blocking=True)
            self._staged.record(side)
            self._collective(self.gathered, self._stage)
            if read:
                self._host[k].copy_(self.gathered, non_blocking=True)
            self._landed[k].record(side)
        main.wait_event(self._staged)  # the next step may overwrite the wire once the snapshot exists
        return k

    def wait(self, ticket: int) -> List[np.ndarray]:
        """Blocks the HOST until the gather (and read) of ``ticket`` has landed; numpy views per rank."""
        if self.cuda:
            self._landed[ticket].synchronize()
        self.host = self._host[ticket]
        return list(self.host.numpy().reshape(self.world, self.words))

    def join(self):
        """Makes the current stream wait for every side-stream gather issued so far (device-side join)."""
        if self.cuda:
            torch.cuda.current_stream(self.wire.device).wait_stream(self._side)


class GradBucket:
    """Data-parallel gradient exchange for the training config (reference: MirroredStrategy's implicit
    all-reduce, code/train.py:55-56, and the loss SUM of code/yolo3/train.py:66-70).

    All gradients live in ONE flat fp32 bucket (2.63 M elements = 10.5 MB for MobileNetV2-0.75 COCO), padded
    to a multiple of the world size (each shard a multiple of 32 elements).  ``reduce_scatter`` leaves each rank with the SUM of its 1/N shard
    (``ncclReduceScatter`` over NVLink; a sharded optimizer updates that shard), ``all_gather`` rebuilds the
    full vector (updated parameters).  The reference sums across replicas (it does not average, each
    replica divides by its LOCAL batch, code/yolo3/model.py:624-625) - so does this."""

    def __init__(self, numel: int, world: int, rank: int, device=None, group=None):
        if numel <= 0 or world <= 0 or not (0 <= rank < world):
            raise ValueError("bad bucket: numel=%d world=%d rank=%d" % (numel, world, rank))
        self.numel, self.world, self.rank, self.group = numel, world, rank, group
        # shards start on 128-byte boundaries (the optimizer kernel uses 128-bit accesses on a rank's shard)
        self.shard_numel = ((numel + world - 1) // world + 31) // 32 * 32
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
