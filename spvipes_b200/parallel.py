"""Data-parallel training: one process per GPU, minibatches sharded across ranks, one gradient all-reduce per step
(the reference is single-device).  BatchNorm statistics, label-rank pairing and the sub-plan
argmax are per-minibatch quantities in the reference, so each rank runs the reference semantics on its own minibatch and
only the parameter gradients are exchanged (NCCL all-reduce over NVLink 5 / NVSwitch; gloo in the CPU tests).

Two all-reduces per step, in backward-completion order: the decoder range of the flat gradient buffer as soon as the decoder
and PoE backward are done (it then overlaps the encoder backward), the encoder range after it; Adam waits for both.
The collectives stay outside CUDA-graph capture: the step is three captured graphs with the two eager all-reduces between.
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
    """sum the gradient buffer over ranks; returns the 1/world factor that the Adam kernel folds into its gradient read
    (no separate scaling pass).  The flat layout is [all encoder blocks | all decoder blocks] (engine.FlatStore phases), so
    each phase is ONE all-reduce: `start(engine, phase)` issues it asynchronously (the decoder range while the encoder
    backward still runs), `finish()` makes the current stream wait for everything issued."""

    def __init__(self, engine, dist, n_buckets: int = 0):
        self.dist = dist
        self.world = dist.get_world_size()
        self.phase_slices = {ph: engine.grads[lo:hi] for ph, (lo, hi) in engine.params.ranges.items()}
        self._work = []

    def start(self, engine, phase):
        self._work.append(self.dist.all_reduce(self.phase_slices[phase], op=self.dist.ReduceOp.SUM, async_op=True))

    def finish(self) -> float:
        for w in self._work:
            w.wait()
        self._work = []
        return 1.0 / self.world

    def __call__(self, engine) -> float:
        for ph in sorted(self.phase_slices, reverse=True):  # decoder range first: complete first in the backward
            self.start(engine, ph)
        return self.finish()


def make_grad_sync(engine, dist):
    """the gradient synchroniser bench.py / the NCCL test use for this process group"""
    return GradSync(engine, dist)


def broadcast_params(engine, dist, src: int = 0):
    """identical initial weights / running statistics on every rank"""
    dist.broadcast(engine.params.flat, src=src)
    dist.broadcast(engine.buffers.flat, src=src)
