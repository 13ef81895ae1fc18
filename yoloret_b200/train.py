"""Training-step pieces of the reference's loop (code/train.py, code/yolo3/train.py) on the B200 engine: the
epoch-wise cosine learning-rate schedule, the Adam update on a sharded flat parameter vector, and checkpoint writing.

  cosine_decay     tf.keras.experimental.CosineDecay(lr0, epochs)(epoch), applied per epoch by a LearningRateScheduler
                   (reference code/train.py:92-100)
  ShardedAdam      tf.keras.optimizers.Adam(lr, epsilon=1e-8) (code/train.py:158-160,195-197) over the data-parallel
                   gradient bucket: reduce-scatter (SUM, like MirroredStrategy) -> Adam on this rank's 1/N shard
                   (yr_adam_step, csrc/optim.cu) -> all-gather of the updated parameters
  save_weights     model.save_weights(path) (code/train.py:74-79,182-186): Keras weights-only HDF5 (h5write.py)
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .parallel import GradBucket


def cosine_decay(lr0: float, decay_steps: int, step: int) -> float:
    """lr0 * 0.5 * (1 + cos(pi * min(step, decay_steps) / decay_steps)), evaluated in float32 like TF."""
    f = np.float32
    s = min(f(step), f(decay_steps))
    frac = f(s / f(decay_steps))
    return float(f(f(lr0) * f(f(0.5) * (f(1.0) + np.cos(f(np.pi) * frac, dtype=f)))))


class ShardedAdam:
    """Adam over a ``GradBucket``: every rank keeps the full fp32 parameter vector, but moments and the update only for
    its own shard (ZeRO-1 style; the reference replicates everything under MirroredStrategy - same arithmetic, 1/N of
    the optimizer traffic per GPU).  ``step()``: reduce-scatter the gradients written into ``bucket.view()``, update the
    shard, all-gather the parameters into ``params``."""

    def __init__(self, params: torch.Tensor, bucket: GradBucket, lr: float, epochs: Optional[int] = None,
                 beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-8):
        if params.dtype != torch.float32 or params.dim() != 1 or params.numel() != bucket.numel:
            raise ValueError("params must be a flat float32 tensor of the bucket's size")
        if not params.is_cuda:
            raise _lib.YrError("ShardedAdam needs CUDA tensors (no CPU fallback exists)")
        self.params, self.bucket = params, bucket
        self.lr0, self.epochs = float(lr), epochs
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations, self.epoch = 0, 0
        n = bucket.shard_numel
        dev = params.device
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        # this rank's parameter shard (padded like the bucket) and the gathered full vector
        self.full = torch.zeros(bucket.padded, dtype=torch.float32, device=dev)
        self.full[:bucket.numel].copy_(params)
        lo = bucket.rank * n
        self.shard = self.full[lo:lo + n]

    @property
    def lr(self) -> float:
        return self.lr0 if self.epochs is None else cosine_decay(self.lr0, self.epochs, self.epoch)

    def on_epoch_begin(self, epoch: int):
        """LearningRateScheduler semantics: the rate of epoch ``epoch`` (0-based) holds for the whole epoch."""
        self.epoch = int(epoch)

    def step(self) -> torch.Tensor:
        b = self.bucket
        g = b.reduce_scatter()
        self.iterations += 1
        with torch.cuda.device(self.params.device):
            _lib.check(_lib.lib().yr_adam_step(self.shard.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                               b.shard_numel, self.lr, self.beta_1, self.beta_2, self.epsilon,
                                               self.iterations, torch.cuda.current_stream(self.params.device).cuda_stream),
                       "yr_adam_step")
        if b.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.full, self.shard.clone(), group=b.group)
        self.params.copy_(self.full[:b.numel])
        return self.params


def save_weights(path: str, weights, layer_order=None):
    """``model.save_weights(path)`` of the reference (code/train.py:74-79,182-186): Keras weights-only HDF5."""
    from .h5write import save_keras_weights
    save_keras_weights(path, weights, layer_order)
