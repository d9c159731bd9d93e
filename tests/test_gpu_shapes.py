"""-m gpu: oracle parity at the EXACT minibatch shapes of BASELINE.json's configs, both precision modes, through the C ABI.

  C2  label PoE,       B  512/group, 5 000 genes, n_hidden 128
  C3  paired OT PoE,   B 1024/group, 10 000 genes, n_hidden 128   (sub-plan gathered from a larger plan with scattered indices)
  C4  cluster OT PoE,  B 2048/group, 20 000 genes, n_hidden 256, 10 clusters
  C5  label PoE,       B 2048/group, 20 000 genes, n_hidden 128

The CPU oracle (oracle/restatement.py, pinned to the unmodified reference in tests/test_oracle.py) takes 0.1-10 s per step at
these sizes; it is evaluated once per config and checked against both modes.  Gates (BASELINE.json north_star): pairing
indices bit-exact; loss and each of the 2 rec + 4 KL terms <= 1e-4 (fp32 mode) / <= 1e-2 (tensor-core mode); latent means,
log-variances and scales <= 1e-3 in BOTH modes (the tensor-core mode runs the K = genes contraction on split-bf16 operands);
gradients per parameter (max |error| / max |gradient|) <= 2e-3 (fp32) / <= GRAD_TOL_TC (tensor-core mode), with the gates of
ReLU units whose pre-activation is zero to within rounding taken from the implementation (asserted to be only those).
The minibatch is gathered with scattered row indices from a 4x larger device-resident count matrix, as the training loop does."""
import functools

import numpy as np
import pytest
import torch

from tests.helpers import grad_errors, relerr
from tests.gpu_helpers import engine_outputs, gate_consistent_grads

pytestmark = pytest.mark.gpu

S, P = 25, 10
CONFIGS = {
    # name: (mode, B, genes, n_hidden, n_labels)
    "C2": ("label", 512, 5000, 128, 10),
    "C3": ("paired", 1024, 10000, 128, 10),
    "C4": ("cluster", 2048, 20000, 256, 10),
    "C5": ("label", 2048, 20000, 128, 10),
    # C4's plan residency: the transport plan stored as bf16 (bench.py keeps the 200k x 200k plan of C4 resident that way);
    # the oracle sees the plan AS STORED, so the gates are the usual ones
    "C4-bf16-plan": ("cluster", 512, 3000, 256, 10),
    "C3-bf16-plan": ("paired", 512, 3000, 128, 10),
}
PLAN_DTYPE = {"C4-bf16-plan": torch.bfloat16, "C3-bf16-plan": torch.bfloat16}
TERMS = ("rec", "kl_private", "kl_poe")
LATENTS = ("private_loc", "private_logvar", "shared_loc", "shared_logvar", "poe_loc", "poe_logvar", "poe_scale")
GRAD_TOL_TC = 2e-3


def _inputs(cfg):
    from spvipes_b200 import synth
    mode, B, G, H, NL = CONFIGS[cfg]
    N = 4 * B
    data = synth.make_counts((N, N), (G, G), NL, device="cuda", seed=31 + len(cfg) * 7 + B)
    gen = torch.Generator().manual_seed(B + G)
    rows = [torch.randperm(N, generator=gen)[:B].to(torch.int32) for _ in (0, 1)]
    eps_p = [torch.randn(B, P, generator=gen) for _ in (0, 1)]
    eps_q = [torch.randn(B, S, generator=gen) for _ in (0, 1)]
    drop = [(torch.rand(B, 2 * H, generator=gen) < 0.9).float() / 0.9 for _ in (0, 1)]
    plan = None
    if mode != "label":
        plan = synth.make_plan(N, N, data.labels[0], data.labels[1], NL, device="cuda", seed=7, dtype=PLAN_DTYPE.get(cfg, torch.float32))
    return data, rows, eps_p, eps_q, drop, plan


@functools.lru_cache(maxsize=1)
def _case(cfg):
    """inputs + initial weights + the oracle's outputs and gradients (fp32, CPU) for one config"""
    from oracle import restatement as rs
    from spvipes_b200.engine import StepEngine
    from spvipes_b200.trainer import init_params
    mode, B, G, H, NL = CONFIGS[cfg]
    data, rows, eps_p, eps_q, drop, plan = _inputs(cfg)
    eng = StepEngine((G, G), H, S, P, 0.1, mode, "cuda", plan=plan, precision="fp32")
    sd0 = init_params(eng, 3)
    del eng
    dm = {(g, k): drop[g][:, i * H:(i + 1) * H] for g in (0, 1) for i, k in enumerate(("private", "shared"))}
    x = [data.X[g].cpu().to(torch.int32)[rows[g].long()].to(torch.float32) for g in (0, 1)]
    labels = [data.labels[g].cpu()[rows[g].long()].numpy() for g in (0, 1)]
    sub = None
    if plan is not None:
        sub = plan[rows[0].long().cuda()][:, rows[1].long().cuda()].float().cpu()
    def oracle(gates=None, probe=None):
        sd = {k: v.clone().requires_grad_("running" not in k) for k, v in sd0.items()}
        out = rs.step(sd, x, mode=mode, n_shared=S, n_private=P, eps_private=eps_p, eps_poe=eps_q, sub=sub,
                      labels=labels if mode in ("label", "cluster") else None, drop_masks=dm, kl_weight=0.25, gates=gates, probe=probe)
        out["loss"].backward()
        return out, {k: sd[k].grad.detach() for k in rs.param_names(sd)}

    probe = {}
    want, grads = oracle(probe=probe)
    want = {k: ([t.detach() if torch.is_tensor(t) else t for t in v] if isinstance(v, (list, tuple)) else (v.detach() if torch.is_tensor(v) else v))
            for k, v in want.items() if k != "new_stats"}
    return (data, rows, eps_p, eps_q, drop, plan), sd0, want, grads, probe, dm, (lambda gates: oracle(gates=gates)[1])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_config_shape_against_oracle(cfg, precision):
    from spvipes_b200.engine import GroupBatch, Noise, StepEngine
    mode, B, G, H, NL = CONFIGS[cfg]
    (data, rows, eps_p, eps_q, drop, plan), sd0, want, grads, probe, dm, rerun = _case(cfg)
    eng = StepEngine((G, G), H, S, P, 0.1, mode, "cuda", plan=plan, precision=precision)
    eng.load_state_dict(sd0)
    eng.set_kl_weight(0.25)
    noise = Noise([e.cuda() for e in eps_p], [e.cuda() for e in eps_q], [d.cuda() for d in drop])
    batches = []
    for g in (0, 1):
        r = rows[g].cuda()
        lab = data.labels[g] if mode in ("label", "cluster") else None
        # labels per cell of the resident matrix, gathered by the kernels with the same row indices; the plan is indexed by
        # the cells' within-group positions = their rows here
        batches.append(GroupBatch(X=data.X[g], rows=r, labels=lab, labels_per_cell=mode == "label", idx=r))
        if mode == "cluster":
            batches[-1].labels = data.labels[g][r.long()].contiguous()
    ws = eng.forward(batches, training=True, noise=noise)
    eng.backward()
    torch.cuda.synchronize()
    out = engine_outputs(eng, ws)
    tol = 1e-4 if precision == "fp32" else 1e-2
    assert relerr(out["loss"], want["loss"]) < tol
    for k in TERMS:
        for g in (0, 1):
            assert relerr(out[k][g], want[k][g]) < tol, (k, g, relerr(out[k][g], want[k][g]))
    for k in LATENTS:
        for g in (0, 1):
            assert relerr(out[k][g], want[k][g]) < 1e-3, (k, g, relerr(out[k][g], want[k][g]))
    if mode in ("label", "paired"):
        for g in (0, 1):
            assert np.array_equal(out["partners"][g], want["partners"][g]), g
    # ReLU units whose pre-activation is ~0 (|pre| < 1e-4 of the layer's largest) have no well-defined gate: on those, and only
    # those (asserted), the oracle's gradient is evaluated with the engine's decision (tests/helpers.py: gate_consistent)
    grads, switched = gate_consistent_grads(eng, ws, probe, grads, rerun, dm)
    worst, where = grad_errors({k: v.cpu() for k, v in eng.grad_dict().items()}, grads)
    assert worst < (2e-3 if precision == "fp32" else GRAD_TOL_TC), (worst, where, switched)
