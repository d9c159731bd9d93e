"""Data-parallel training: one process per GPU, minibatches sharded across ranks, one gradient all-reduce per step
(the reference is single-device).  BatchNorm statistics, label-rank pairing and the sub-plan
argmax are per-minibatch quantities in the reference, so each rank runs the reference semantics on its own minibatch and
only the parameter gradients are exchanged (NCCL all-reduce over NVLink 5 / NVSwitch; gloo in the CPU tests).

Two all-reduces per step, in backward-completion order: the decoder range of the flat gradient buffer as soon as the decoder
and PoE backward are done (it then overlaps the encoder backward), the encoder range after it; Adam waits for both.
Product path on NVLink boxes: `NvlinkGradSync` - the gradient buffer lives in symmetric memory and each range is summed by ONE
hand-written kernel (csrc/xgpu.cu: multimem.ld_reduce / multimem.st through the NVSwitch, or peer loads / stores) that is
captured INSIDE the step's CUDA graph: one graph launch per step, no host in the loop.  `GradSync` (torch.distributed
all-reduce issued eagerly between four captured graphs) remains for gloo (CPU tests) and as the A/B baseline (SPV_DP_SYNC=nccl).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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

    in_graph = False
    kind = "torch.distributed all-reduce (eager, between four graphs)"

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


class NvlinkGradSync:
    """in-graph gradient all-reduce over NVLink (csrc/xgpu.cu).  Construction is collective: every rank allocates the flat
    gradient buffer and a flag buffer in symmetric memory (torch.distributed._symmetric_memory: allocation + exchange of the
    peer / multicast mappings only - the reduction itself is this library's kernel) and the engine adopts the buffer.
    `allreduce(phase, channel)` enqueues the kernel for one phase of the flat layout on the current stream."""

    in_graph = True

    def __init__(self, engine, dist, multicast: Optional[bool] = None, blocks: int = 0):
        import os
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.dist, self.engine, self.lib = dist, engine, L.load()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        group = dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass  # newer torch enables it implicitly
        dev = engine.device
        n = engine.params.numel
        self.buf = symm.empty(n, dtype=torch.float32, device=dev)
        self.flags = symm.empty(self.lib.spv_xgpu_flag_ints(), dtype=torch.int32, device=dev)
        self.buf.zero_()
        self.flags.zero_()
        hb = symm.rendezvous(self.buf, group.group_name)
        hf = symm.rendezvous(self.flags, group.group_name)
        torch.cuda.synchronize(dev)
        dist.barrier()  # every rank's flags are zero before anybody's first kernel signals
        self._handles = (hb, hf)
        self.peer_bufs = L.ptr_array([int(x) for x in hb.buffer_ptrs])
        self.peer_flags = L.ptr_array([int(x) for x in hf.buffer_ptrs])
        mc = int(getattr(hb, "multicast_ptr", 0) or 0)
        if multicast is None:
            multicast = os.environ.get("SPV_DP_MULTICAST", "1") == "1"
        self.mc = mc if (multicast and mc) else None
        self.kind = "nvlink-multimem" if self.mc else "nvlink-p2p"
        self.blocks = blocks or int(os.environ.get("SPV_DP_BLOCKS", "0"))
        self.state = torch.zeros(12, dtype=torch.int32, device=dev)
        engine.grads = self.buf  # the backward kernels now write the symmetric buffer directly
        self.ranges = dict(engine.params.ranges)

    def allreduce(self, phase: int, channel: int, blocks: int = 0):
        self.allreduce_range(*self.ranges[phase], channel, blocks)

    def allreduce_range(self, lo: int, hi: int, channel: int, blocks: int = 0):
        """[lo, hi) of the flat gradient buffer (multiples of 4 floats: every block of the layout is 16-byte aligned).
        blocks: CTAs of the exchange kernel (0: the synchroniser's default) - few when it runs beside compute it must not crowd out"""
        from . import _lib as L
        L.check(self.lib.spv_xgpu_allreduce(self.peer_bufs, self.peer_flags, self.mc, lo, hi - lo, self.rank, self.world, channel,
                                            L.ptr(self.state), blocks or self.blocks, torch.cuda.current_stream(self.engine.device).cuda_stream),
                "spv_xgpu_allreduce")

    def check(self):
        """raise if a handshake of any launch so far timed out (synchronises the device)"""
        bad = int(self.state[8].item())
        if bad:
            raise RuntimeError(f"NVLink gradient all-reduce: rank {bad - 1} did not answer rank {self.rank} within the time limit")

    def __call__(self, engine) -> float:
        """all ranges, on the current stream (eager use: tests)"""
        for ch, ph in enumerate(sorted(self.ranges, reverse=True)):
            self.allreduce(ph, ch)
        return 1.0 / self.world


def make_grad_sync(engine, dist):
    """the gradient synchroniser bench.py / the NCCL tests use for this process group: the in-graph NVLink all-reduce on CUDA
    with the NCCL backend, torch.distributed all-reduces otherwise (gloo) or when SPV_DP_SYNC=nccl asks for the baseline.
    If symmetric memory cannot be set up on this box the NCCL path is used and the reason is kept in `.fallback_reason`."""
    import os
    want = os.environ.get("SPV_DP_SYNC", "nvlink")
    if want == "nvlink" and engine.device.type == "cuda" and dist.get_backend() == "nccl":
        try:
            return NvlinkGradSync(engine, dist)
        except Exception as e:  # no peer access / no symmetric-memory support: keep training, say why
            gs = GradSync(engine, dist)
            gs.fallback_reason = f"{type(e).__name__}: {e}"[:200]
            return gs
    return GradSync(engine, dist)


def broadcast_params(engine, dist, src: int = 0):
    """identical initial weights / running statistics on every rank"""
    dist.broadcast(engine.params.flat, src=src)
    dist.broadcast(engine.buffers.flat, src=src)
