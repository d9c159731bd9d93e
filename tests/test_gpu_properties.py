"""-m gpu: size-independent properties of the hot path at BASELINE.json's configs[1] minibatch shape (2 x 512 cells, 5000
genes, H 128, label PoE), where the CPU oracle is too slow to be the checker:
  * determinism: the same seeded step twice -> bit-identical loss terms, gradients and updated parameters;
  * CUDA-graph replay == eager launches of the same step (bitwise);
  * the order of the genes is irrelevant: permuting the count columns together with the gene-indexed parameters leaves
    loss terms and latents unchanged (up to summation order);
  * the training step makes progress: the loss falls over 30 optimiser steps on a fixed minibatch stream.
All through the C ABI (spvipes_b200.engine / trainer)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

B, G, H, S, P, NL = 512, 5000, 128, 25, 10, 10


def _setup(precision="bf16", seed=3, n_cells=2048):
    from spvipes_b200 import synth
    from spvipes_b200.engine import StepEngine
    from spvipes_b200.trainer import init_params
    data = synth.make_counts((n_cells, n_cells), (G, G), NL, device="cuda", seed=2024)
    eng = StepEngine((G, G), H, S, P, 0.1, "label", "cuda", seed=11, precision=precision)
    init_params(eng, seed)
    return data, eng


def _batches(data, step):
    from spvipes_b200.engine import GroupBatch
    gen = torch.Generator(device="cuda").manual_seed(100 + step)
    rows = [torch.randperm(data.X[g].shape[0], generator=gen, device="cuda")[:B].to(torch.int32) for g in (0, 1)]
    return [GroupBatch(X=data.X[g], rows=rows[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_step_is_deterministic(precision):
    from spvipes_b200.trainer import TrainLoop
    out = []
    for _ in range(2):
        data, eng = _setup(precision)
        loop = TrainLoop(eng)
        loop.set_epoch(10)
        for s in range(3):
            loop.step(_batches(data, s))
        torch.cuda.synchronize()
        out.append((eng.loss_out.clone(), eng.grads.clone(), eng.params.flat.clone(), eng.buffers.flat.clone()))
    for a, b in zip(*out):
        assert torch.equal(a, b)


def test_graph_replay_equals_eager_step():
    from spvipes_b200.engine import GroupBatch
    from spvipes_b200.trainer import TrainLoop
    res = []
    for use_graph in (False, True):
        data, eng = _setup("bf16")
        loop = TrainLoop(eng)
        loop.set_epoch(10)
        rows_cur = [torch.zeros(B, dtype=torch.int32, device="cuda") for _ in (0, 1)]
        static = [GroupBatch(X=data.X[g], rows=rows_cur[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
        graph = None
        if use_graph:
            for g in (0, 1):
                rows_cur[g].copy_(_batches(data, 0)[g].rows)
            graph = loop.capture(static)  # restores parameters, moments, running statistics and the step counter
        for s in range(4):
            bt = _batches(data, s)
            for g in (0, 1):
                rows_cur[g].copy_(bt[g].rows)
            if graph is not None:
                graph.replay()
            else:
                loop.step(static)
        torch.cuda.synchronize()
        res.append((eng.loss_out.clone(), eng.params.flat.clone(), eng.adam_m.clone(), eng.buffers.flat.clone(), int(eng.step_dev)))
    for a, b in zip(*res):
        assert torch.equal(a, b) if torch.is_tensor(a) else a == b


def test_gene_order_is_irrelevant():
    from spvipes_b200.engine import GroupBatch, Noise
    data, eng = _setup("fp32")
    gen = torch.Generator().manual_seed(5)
    noise = Noise([torch.randn(B, P, generator=gen).cuda() for _ in (0, 1)], [torch.randn(B, S, generator=gen).cuda() for _ in (0, 1)],
                  [((torch.rand(B, 2 * H, generator=gen) < 0.9).float() / 0.9).cuda() for _ in (0, 1)])
    X = [data.X[g][:B].contiguous() for g in (0, 1)]
    lab = [data.labels[g][:B].contiguous() for g in (0, 1)]
    ws = eng.forward([GroupBatch(X=X[g], labels=lab[g]) for g in (0, 1)], training=True, noise=noise)
    torch.cuda.synchronize()
    ref = (eng.loss_out.clone(), [w.rec.clone() for w in ws], [w.zpoe.clone() for w in ws], [w.zpriv.clone() for w in ws])
    sd = eng.state_dict()
    perm = [torch.randperm(G, generator=gen) for _ in (0, 1)]
    sd2 = {}
    for k, v in sd.items():
        g = 0 if ("encoder_0" in k or "decoder_0" in k or k.endswith("px_r.0")) else 1
        pg = perm[g].to(v.device)
        if k.endswith("fc1.weight"):
            sd2[k] = v[:, pg]                                  # [H, G]: gene = input column
        elif ("decoder_" in k or k.startswith("px_r")) and v.dim() >= 1 and v.shape[0] == G:
            sd2[k] = v[pg]                                      # gene-indexed rows (regressor / mixture weights, BN vectors, px_r)
        else:
            sd2[k] = v
    data2, eng2 = _setup("fp32")
    eng2.load_state_dict(sd2)
    X2 = [X[g].view(torch.int16)[:, perm[g].cuda()].contiguous().view(torch.uint16) for g in (0, 1)]
    ws2 = eng2.forward([GroupBatch(X=X2[g], labels=lab[g]) for g in (0, 1)], training=True, noise=noise)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert rel(eng2.loss_out[:7], ref[0][:7]) < 2e-5
    for g in (0, 1):
        assert rel(ws2[g].rec, ref[1][g]) < 2e-5
        assert rel(ws2[g].zpoe, ref[2][g]) < 1e-4 and rel(ws2[g].zpriv, ref[3][g]) < 1e-4


def test_training_makes_progress():
    from spvipes_b200.trainer import TrainLoop
    data, eng = _setup("bf16")
    loop = TrainLoop(eng, n_epochs_kl_warmup=None)
    loop.set_epoch(1)
    losses = []
    for s in range(30):
        loop.step(_batches(data, s % 4))
        losses.append(float(eng.loss_out[0]))
    assert all(l == l for l in losses)  # no NaN
    assert sum(losses[-5:]) / 5 < 0.97 * sum(losses[:5]) / 5, losses
