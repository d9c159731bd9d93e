"""-m "not gpu": host-side logic of the data-parallel path with world_size 2 over gloo on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeEngine:
    def __init__(self, n, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self.grads = torch.randn(n, generator=g)

        class _P:
            pass
        self.params, self.buffers = _P(), _P()
        self.params.ranges = {0: (0, 400), 1: (400, n)}  # [encoder range | decoder range] of the flat layout
        self.params.flat = torch.full((8,), float(rank))
        self.buffers.flat = torch.full((4,), float(rank) + 0.5)


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spvipes_b200.parallel import GradSync, broadcast_params, shard_rows
    eng = _FakeEngine(n, rank)
    sync = GradSync(eng, dist)
    sync.start(eng, 1)  # the way the step uses it: decoder range first (asynchronously), then the encoder range
    sync.start(eng, 0)
    scale = sync.finish()
    again = GradSync(eng, dist)(eng)  # one-shot form: sums once more
    eng.grads /= 2.0  # two ranks: every entry was summed over identical (already reduced) buffers -> doubled
    assert again == scale
    broadcast_params(eng, dist, src=0)
    rows = shard_rows(np.arange(1001), rank, world)
    if rank == 0:
        torch.save({"grads": eng.grads * scale, "scale": scale}, out)
    assert float(eng.params.flat[0]) == 0.0 and float(eng.buffers.flat[0]) == 0.5
    assert len(rows) == 500 and rows[0] == rank * 500
    dist.destroy_process_group()


def test_gradsync_world2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = 1003
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = torch.load(out)
    want = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in (0, 1)) / 2
    assert got["scale"] == 0.5
    assert torch.allclose(got["grads"], want, atol=1e-6)


def test_bucket_bounds_cover_buffer():
    from spvipes_b200.parallel import bucket_bounds
    for n, k in ((1003, 3), (16, 4), (5, 8)):
        b = bucket_bounds(n, k)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
