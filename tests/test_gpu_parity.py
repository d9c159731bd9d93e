"""-m gpu: the CUDA path (through the C ABI) against the golden vectors generated from the unmodified reference and
against the oracle (oracle/restatement.py) on the same seeded inputs.

Tolerances (BASELINE.json north_star, fp32 mode): indices bit-exact; loss and each of the 2 reconstruction + 4 KL terms
<= 1e-4 relative; latent means / variances <= 1e-3; gradients <= 2e-3 (per parameter, normalised by its max |grad|)."""
import numpy as np
import pytest
import torch

from tests.helpers import Golden, golden_names, grad_errors, relerr, run_oracle
from tests.gpu_helpers import engine_from_golden, engine_outputs

pytestmark = pytest.mark.gpu

TERMS_1E4 = ("rec", "kl_private", "kl_poe", "library")
LATENTS_1E3 = ("private_loc", "private_logvar", "private_log_z", "shared_loc", "shared_logvar", "poe_loc", "poe_logvar",
               "poe_scale", "poe_log_z")


@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_golden(name):
    gd = Golden(name)
    eng, batches, noise = engine_from_golden(gd)
    ws = eng.forward(batches, training=gd.training, noise=noise)
    torch.cuda.synchronize()
    out = engine_outputs(eng, ws)
    assert relerr(out["loss"], gd.out["loss"]) < 1e-4
    for k in TERMS_1E4:
        for g in (0, 1):
            assert relerr(out[k][g].reshape(-1), gd.out[f"{k}{g}"].reshape(-1)) < 1e-4, (k, g)
    for k in LATENTS_1E3:
        for g in (0, 1):
            assert relerr(out[k][g], gd.out[f"{k}{g}"]) < 1e-3, (k, g)
    if gd.mode in ("label", "paired"):
        for g in (0, 1):  # integer pairing contract: bit-exact
            assert np.array_equal(out["partners"][g], gd.out[f"partner{g}"]), g
    if gd.training:  # BatchNorm running statistics after the step
        sd = eng.state_dict()
        for k, v in gd.after.items():
            assert relerr(sd[k].cpu(), v) < 1e-4, k


@pytest.mark.parametrize("name", golden_names(training=True))
def test_backward_matches_golden(name):
    gd = Golden(name)
    eng, batches, noise = engine_from_golden(gd)
    eng.forward(batches, training=True, noise=noise)
    eng.backward()
    torch.cuda.synchronize()
    got = {k: v.cpu() for k, v in eng.grad_dict().items()}
    worst, where = grad_errors(got, gd.grads)
    assert worst < 2e-3, (worst, where)


@pytest.mark.parametrize("name", ["label_tiny", "paired_tiny", "cluster_tiny"])
def test_matches_oracle_f64(name):
    """against the float64 oracle: the fp32 CUDA path is as close to the exact answer as the fp32 reference is"""
    gd = Golden(name)
    o64, g64, _ = run_oracle(gd, torch.float64)
    eng, batches, noise = engine_from_golden(gd)
    ws = eng.forward(batches, training=True, noise=noise)
    eng.backward()
    torch.cuda.synchronize()
    out = engine_outputs(eng, ws)
    assert relerr(out["loss"], o64["loss"]) < 2e-5
    for k in ("rec", "kl_private", "kl_poe"):
        for g in (0, 1):
            assert relerr(out[k][g], o64[k][g]) < 5e-5, (k, g)
    worst, where = grad_errors({k: v.cpu() for k, v in eng.grad_dict().items()}, {k: v.detach() for k, v in g64.items()})
    assert worst < 2e-3, (worst, where)


def test_adam_step_matches_oracle():
    from oracle import restatement as rs
    gd = Golden("label_tiny")
    eng, batches, noise = engine_from_golden(gd)
    p0 = eng.params.flat.clone()
    m = torch.zeros_like(p0); v = torch.zeros_like(p0)
    p = p0.clone()
    for t in (1, 2, 3):
        eng.forward(batches, training=True, noise=noise)
        eng.backward()
        g = eng.grads.clone()
        p, m, v = rs.adam_step(p.double(), g.double(), m.double(), v.double(), t)
        eng.adam_step()
        torch.cuda.synchronize()
        assert relerr(eng.params.flat, p) < 1e-5
        assert int(eng.step_dev.item()) == t
        p, m, v = eng.params.flat.clone(), eng.adam_m.clone(), eng.adam_v.clone()
