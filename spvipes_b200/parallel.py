"""Data-parallel training: one process per GPU, minibatches sharded across ranks, one gradient all-reduce per step
(the reference is single-device).  BatchNorm statistics, label-rank pairing and the sub-plan
argmax are per-minibatch quantities in the reference, so each rank runs the reference semantics on its own minibatch and
only the parameter gradients are exchanged (NCCL all-reduce over NVLink 5 / NVSwitch; gloo in the CPU tests).

The all-reduce is issued on the step's stream as a few large buckets in backward-completion order (decoder of group 1,
decoder of group 0, encoders) so that NCCL pipelines them; inside the captured CUDA graph the buckets become graph nodes
that overlap the tail of the backward kernels of the other group's stream.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def shard_rows(group_indices: Sequence[int], rank: int, world: int) -> np.ndarray:
    """contiguous 1/world shard of a group's (already permuted) training rows; every rank gets the same count so that all
    ranks run the same number of steps (the remainder rows are dropped, like drop_last)"""
    idx = np.asarray(group_indices)
    per = len(idx) // world
    return idx[rank * per:(rank + 1) * per]


def bucket_bounds(numel: int, n_buckets: int) -> List[Tuple[int, int]]:
    per = (numel + n_buckets - 1) // n_buckets
    per = (per + 3) // 4 * 4
    return [(i, min(numel, i + per)) for i in range(0, numel, per)]


class GradSync:
    """sum the flat gradient buffer over ranks in `n_buckets` chunks; returns the 1/world factor that the fused Adam kernel
    folds into its gradient read (no separate scaling pass)."""

    def __init__(self, engine, dist, n_buckets: int = 0):
        self.dist = dist
        if n_buckets <= 0:  # one bucket up to 64 MiB of gradients (latency bound on NVSwitch), then 32 MiB buckets
            n_buckets = max(1, (engine.grads.numel() * 4 + (64 << 20) - 1) // (64 << 20) * 2 - 1)
        self.world = dist.get_world_size()
        self.buckets = [engine.grads[a:b] for a, b in reversed(bucket_bounds(engine.grads.numel(), n_buckets))]

    def __call__(self, engine) -> float:
        for b in self.buckets:
            self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM)
        return 1.0 / self.world


def broadcast_params(engine, dist, src: int = 0):
    """identical initial weights / running statistics on every rank"""
    dist.broadcast(engine.params.flat, src=src)
    dist.broadcast(engine.buffers.flat, src=src)
