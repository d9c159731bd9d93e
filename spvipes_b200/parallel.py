"""Data-parallel training: one process per GPU, minibatches sharded across ranks, one gradient all-reduce per step
(the reference is single-device).  BatchNorm statistics, label-rank pairing and the sub-plan
argmax are per-minibatch quantities in the reference, so each rank runs the reference semantics on its own minibatch and
only the parameter gradients are exchanged (NCCL all-reduce over NVLink 5 / NVSwitch; gloo in the CPU tests).
"""
from __future__ import annotations

import torch


class GradSync:
    """average the flat gradient buffer over ranks, in `n_buckets` chunks issued back to back on the step's stream."""

    def __init__(self, engine, dist, n_buckets: int = 4):
        self.dist = dist
        self.world = dist.get_world_size()
        n = engine.grads.numel()
        per = (n + n_buckets - 1) // n_buckets
        self.buckets = [engine.grads[i:min(n, i + per)] for i in range(0, n, per)]

    def __call__(self, engine):
        for b in self.buckets:
            self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM)
        return 1.0 / self.world  # Adam folds the 1/world into its gradient read


def broadcast_params(engine, dist, src: int = 0):
    dist.broadcast(engine.params.flat, src=src)
    dist.broadcast(engine.buffers.flat, src=src)
