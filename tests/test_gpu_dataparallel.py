"""-m gpu: the data-parallel step (two ranks, gloo, both on cuda:0 so that a single-GPU box can run it): after two steps
every rank holds the same parameters, and they equal a single process that averages the two ranks' gradients itself."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B, G, H, S, P, NL = 128, 600, 64, 25, 10, 5


def _engine_and_data(rank):
    from spvipes_b200 import synth
    from spvipes_b200.engine import GroupBatch, StepEngine
    from spvipes_b200.trainer import init_params
    data = synth.make_counts((B, B), (G, G), NL, device="cuda", seed=500 + rank)
    eng = StepEngine((G, G), H, S, P, 0.0, "label", "cuda", seed=0, precision="fp32")  # no dropout / fixed noise seed per rank
    init_params(eng, 1)
    batches = [GroupBatch(X=data.X[g], labels=data.labels[g]) for g in (0, 1)]
    return eng, batches


def _noise(rank):
    from spvipes_b200.engine import Noise
    gen = torch.Generator().manual_seed(900 + rank)
    return Noise([torch.randn(B, P, generator=gen).cuda() for _ in (0, 1)], [torch.randn(B, S, generator=gen).cuda() for _ in (0, 1)], None)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spvipes_b200.parallel import GradSync, broadcast_params
    from spvipes_b200.trainer import TrainLoop
    eng, batches = _engine_and_data(rank)
    broadcast_params(eng, dist, src=0)
    loop = TrainLoop(eng)
    loop.grad_sync = GradSync(eng, dist)
    loop.set_epoch(5)
    for _ in range(2):
        loop.step(batches, _noise(rank))
    torch.cuda.synchronize()
    torch.save({"params": eng.params.flat.cpu(), "step": int(eng.step_dev)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_ranks_equal_manual_gradient_average(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [torch.load(str(tmp_path / f"r{r}.pt")) for r in (0, 1)]
    assert torch.equal(got[0]["params"], got[1]["params"]) and got[0]["step"] == got[1]["step"] == 2
    # single process: two engines (one per rank's data) kept in lock step by averaging their gradients by hand
    from spvipes_b200.trainer import TrainLoop
    engs, bts = zip(*[_engine_and_data(r) for r in (0, 1)])
    loops = [TrainLoop(e) for e in engs]
    for lp in loops:
        lp.set_epoch(5)
    for _ in range(2):
        for r in (0, 1):
            engs[r].forward(bts[r], training=True, noise=_noise(r))
            engs[r].backward()
        torch.cuda.synchronize()
        avg = engs[0].grads + engs[1].grads  # summed; Adam folds the 1/world factor
        for r in (0, 1):
            engs[r].grads.copy_(avg)
            engs[r].adam_step(lr=loops[r].lr, eps=loops[r].eps, weight_decay=loops[r].weight_decay, grad_scale=0.5)
    torch.cuda.synchronize()
    want = engs[0].params.flat.cpu()
    assert float((got[0]["params"] - want).abs().max()) <= 1e-6 * float(want.abs().max()) + 1e-8
