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


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["label_batch3_tiny", "paired_batch2_eval"])
def test_batch_covariates_match_reference(name, precision):
    """n_batch > 1: the one-hot batch code appended to the input of the encoders' first layer and of the four decoder nets
    (reference nn/networks.py:60-68, 110-118; scvi FCLayers inject_covariates), against fixtures from the unmodified reference
    (tests/golden_next/, oracle/make_golden.py) - forward gates of the mode, gradients per parameter <= 2e-3."""
    from tests.helpers import GOLDEN_NEXT_DIR
    from tests.gpu_helpers import gate_consistent_grads
    gd = Golden(name, GOLDEN_NEXT_DIR)
    assert gd.n_batch > 1
    eng, batches, noise = engine_from_golden(gd, precision=precision)
    ws = eng.forward(batches, training=gd.training, noise=noise)
    if gd.training:
        eng.backward()
    torch.cuda.synchronize()
    out = engine_outputs(eng, ws)
    tol = 1e-4 if precision == "fp32" else 1e-2
    assert relerr(out["loss"], gd.out["loss"]) < tol
    for k in ("rec", "kl_private", "kl_poe"):
        for g in (0, 1):
            assert relerr(out[k][g].reshape(-1), gd.out[f"{k}{g}"].reshape(-1)) < tol, (k, g)
    for k in LATENTS_1E3:
        for g in (0, 1):
            assert relerr(out[k][g], gd.out[f"{k}{g}"]) < 1e-3, (k, g)
    if gd.training:
        probe = {}
        run_oracle(gd, backward=False, probe=probe)
        want, switched = gate_consistent_grads(eng, ws, probe, gd.grads, lambda gates: run_oracle(gd, gates=gates)[1], gd.drop_masks())
        worst, where = grad_errors({k: v.cpu() for k, v in eng.grad_dict().items()}, want)
        assert worst < 2e-3, (worst, where, switched)
        sd = eng.state_dict()
        for k, v in gd.after.items():  # BatchNorm running statistics after the step
            assert relerr(sd[k].cpu(), v) < (1e-4 if precision == "fp32" else 1e-3), k


def test_get_loadings_strips_covariate_columns():
    """reference module/spVIPESmodule.py:773-807: loadings = diag(gamma / sqrt(running_var + eps)) W, covariate columns dropped"""
    import numpy as np
    from tests.helpers import GOLDEN_NEXT_DIR
    from spvipes_b200.module import spVIPESmodule
    gd = Golden("label_batch3_tiny", GOLDEN_NEXT_DIR)
    m = spVIPESmodule(groups_lengths={0: gd.G0, 1: gd.G1}, groups_obs_names=[None, None], groups_var_names={0: None, 1: None},
                      groups_obs_indices=[None, None], groups_var_indices=[np.arange(gd.G0), np.arange(gd.G0, gd.G0 + gd.G1)],
                      use_labels=True, n_labels=gd.n_labels, n_batch=gd.n_batch, n_hidden=gd.H, n_dimensions_shared=gd.S,
                      n_dimensions_private=gd.P, dropout_rate=gd.dropout)
    m.load_state_dict(gd.sd, strict=True)
    for g, G in enumerate((gd.G0, gd.G1)):
        for kind, dim in (("shared", gd.S), ("private", gd.P)):
            k = f"decoder_{g}.factor_regressor_{kind}.fc_layers.Layer 0"
            w, gamma, rv = gd.sd[k + ".0.weight"], gd.sd[k + ".1.weight"], gd.sd[k + ".1.running_var"]
            want = ((gamma / torch.sqrt(rv + 1e-3)).unsqueeze(1) * w)[:, :-gd.n_batch].numpy()
            got = m.get_loadings(g, kind)
            assert got.shape == (G, dim)
            assert np.allclose(got, want, rtol=1e-6, atol=1e-7)
