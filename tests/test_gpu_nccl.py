"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`): the data-parallel step exactly as bench.py
times it at N > 1 - NCCL process group, one rank per GPU, the CAPTURED step (`TrainLoop.capture` with a gradient sync) replayed
for several steps with in-kernel Philox noise, tensor-core mode.  Checks: every rank ends with bitwise-identical parameters and
optimiser state, and they equal a single process that runs the two ranks' steps itself and averages their gradients by hand.
Three gradient synchronisers: the in-graph NVLink all-reduce kernel (csrc/xgpu.cu) through the switch's multicast reduction,
the same kernel on peer loads / stores, and torch.distributed all-reduces issued between four graphs (the r1 baseline).
A second test drives the all-reduce kernel alone on ranges of awkward sizes."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")]

B, G, H, S, P, NL, STEPS = 256, 1500, 128, 25, 10, 6, 6


def _engine_and_data(rank, dev, precision="bf16"):
    from spvipes_b200 import synth
    from spvipes_b200.engine import GroupBatch, StepEngine
    from spvipes_b200.trainer import init_params
    data = synth.make_counts((4 * B, 4 * B), (G, G), NL, device=dev, seed=500 + rank)
    eng = StepEngine((G, G), H, S, P, 0.1, "label", dev, seed=1000 + rank, precision=precision)
    init_params(eng, 1)
    rows = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in (0, 1)]
    batches = [GroupBatch(X=data.X[g], rows=rows[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
    return eng, batches, rows


def _rows_for(rank, step, dev):
    gen = torch.Generator().manual_seed(77 + 13 * rank + step)
    return [torch.randperm(4 * B, generator=gen)[:B].to(torch.int32).to(dev) for _ in (0, 1)]


def _worker(rank, world, port, out_dir, sync):
    import torch.distributed as dist
    os.environ["SPV_DP_SYNC"] = "nccl" if sync == "nccl" else "nvlink"
    os.environ["SPV_DP_MULTICAST"] = "0" if sync == "nvlink-p2p" else "1"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from spvipes_b200.parallel import make_grad_sync, broadcast_params
    from spvipes_b200.trainer import TrainLoop
    eng, batches, rows = _engine_and_data(rank, dev)
    broadcast_params(eng, dist, src=0)
    loop = TrainLoop(eng)
    loop.grad_sync = make_grad_sync(eng, dist)
    loop.set_epoch(100)
    for g, r in enumerate(_rows_for(rank, 0, dev)):
        rows[g].copy_(r)
    graph = loop.capture(batches)
    for s in range(STEPS):
        for g, r in enumerate(_rows_for(rank, s, dev)):
            rows[g].copy_(r)
        graph.replay()
    torch.cuda.synchronize()
    if hasattr(loop.grad_sync, "check"):
        loop.grad_sync.check()
    torch.save({"params": eng.params.flat.cpu(), "m": eng.adam_m.cpu(), "v": eng.adam_v.cpu(), "step": int(eng.step_dev),
                "loss": eng.loss_out.cpu(), "sync": f"{type(loop.grad_sync).__name__} ({loop.grad_sync.kind})",
                "fallback": getattr(loop.grad_sync, "fallback_reason", None)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(600)
@pytest.mark.parametrize("sync", ["nvlink", "nvlink-p2p", "nccl"])
def test_nccl_captured_step_equals_manual_gradient_average(tmp_path, sync):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), sync), nprocs=2, join=True)
    got = [torch.load(str(tmp_path / f"r{r}.pt")) for r in (0, 1)]
    assert got[0]["fallback"] is None, got[0]["fallback"]  # the synchroniser asked for is the one that ran
    if sync != "nccl":
        assert "NvlinkGradSync" in got[0]["sync"] and ("multimem" in got[0]["sync"]) == (sync == "nvlink"), got[0]["sync"]
    for k in ("params", "m", "v"):
        assert torch.equal(got[0][k], got[1][k]), k  # bitwise across ranks
    assert got[0]["step"] == got[1]["step"] == STEPS
    # single process, one GPU: the two ranks' engines kept in lock step by averaging their gradients by hand
    from spvipes_b200.trainer import TrainLoop
    dev = torch.device("cuda", 0)
    engs, bts, rws = zip(*[_engine_and_data(r, dev) for r in (0, 1)])
    loops = [TrainLoop(e) for e in engs]
    for lp in loops:
        lp.set_epoch(100)
    for e in engs:
        e.stage_in_adam = False  # eager reference: bf16 operand copies refreshed by conversion launches
    for s_ in range(STEPS):
        for r in (0, 1):
            for g, rr in enumerate(_rows_for(r, s_, dev)):
                rws[r][g].copy_(rr)
            engs[r].forward(bts[r], training=True)
            engs[r].backward()
        torch.cuda.synchronize()
        avg = engs[0].grads + engs[1].grads  # summed; Adam folds the 1/world factor
        for r in (0, 1):
            engs[r].grads.copy_(avg)
            engs[r].adam_step(lr=loops[r].lr, eps=loops[r].eps, weight_decay=loops[r].weight_decay, grad_scale=0.5)
    torch.cuda.synchronize()
    want = engs[0].params.flat.cpu()
    err = float((got[0]["params"] - want).abs().max()) / float(want.abs().max())
    print("sync:", got[0]["sync"], "max rel param diff vs manual average:", err)
    assert err <= 2e-6, err


def _allreduce_worker(rank, world, port, out_dir, multicast):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from spvipes_b200.parallel import NvlinkGradSync

    class _Store:  # the two attributes NvlinkGradSync reads from an engine
        pass
    eng = _Store()
    eng.device = dev
    eng.params = _Store()
    # ranges of awkward sizes: one float4, fewer float4s than ranks x CTAs, a size that does not divide by the world size
    eng.params.ranges = {0: (0, 4), 1: (4, 4 + 4 * 3), 2: (16, 16 + 4 * 100_003), 3: (16 + 4 * 100_003, 16 + 4 * 100_003 + 4 * 3_000_000)}
    eng.params.numel = eng.params.ranges[3][1]
    gs = NvlinkGradSync(eng, dist, multicast=multicast)
    ok = True
    for it in range(3):  # repeated launches on the same channels: the epochs advance
        gen = torch.Generator(device=dev).manual_seed(100 * it + rank)
        mine = torch.randn(eng.params.numel, generator=gen, device=dev)
        theirs = [torch.randn(eng.params.numel, generator=torch.Generator(device=dev).manual_seed(100 * it + r), device=dev) for r in range(world)]
        want = theirs[0].clone()
        for r in range(1, world):
            want += theirs[r]
        eng.grads.copy_(mine)
        torch.cuda.synchronize()
        dist.barrier()
        for ch, ph in enumerate(sorted(eng.params.ranges)):
            gs.allreduce(ph, ch)
        torch.cuda.synchronize()
        gs.check()
        ok = ok and bool(torch.equal(eng.grads, want)) if world == 2 else ok and bool(torch.allclose(eng.grads, want, rtol=1e-6, atol=1e-6))
        dist.barrier()
    torch.save({"ok": ok, "kind": gs.kind, "sum": eng.grads.double().sum().item()}, os.path.join(out_dir, f"a{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("multicast", [True, False])
def test_xgpu_allreduce_kernel(tmp_path, multicast):
    """spv_xgpu_allreduce alone: sums equal the plain sum (two ranks: one addition, so bit-exact), identical on both ranks"""
    world = 2
    mp.spawn(_allreduce_worker, args=(world, _free_port(), str(tmp_path), multicast), nprocs=world, join=True)
    got = [torch.load(str(tmp_path / f"a{r}.pt")) for r in range(world)]
    assert all(g["ok"] for g in got), got
    assert got[0]["sum"] == got[1]["sum"]
    assert ("multimem" in got[0]["kind"]) == multicast, got[0]["kind"]
