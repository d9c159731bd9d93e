"""-m "not gpu": the oracle (oracle/restatement.py) against the committed golden vectors
(generated from the unmodified reference, oracle/make_golden.py) and, when the reference
tree is present (authoring container), against the reference itself on fresh inputs."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import restatement as rs
from tests.helpers import GOLDEN_NEXT_DIR, Golden, golden_names, golden_next_names, grad_errors, relerr, run_oracle

TERMS = ("rec", "kl_private", "kl_poe", "library", "private_loc", "private_logvar", "private_log_z",
         "shared_loc", "shared_logvar", "poe_loc", "poe_logvar", "poe_scale", "poe_log_z")


def test_golden_next_fixtures_present():
    assert golden_next_names() == ["label_batch3_tiny", "paired_batch2_eval"]


@pytest.mark.parametrize("name", golden_names() + ["next:" + n for n in golden_next_names()])
def test_oracle_matches_golden(name):
    """every committed fixture of the unmodified reference; the `next:` ones exercise batch covariates (n_batch > 1,
    SURVEY.md 8f rank 3), which only the oracle implements so far"""
    gd = Golden(name[5:], GOLDEN_NEXT_DIR) if name.startswith("next:") else Golden(name)
    out, grads, _ = run_oracle(gd)
    assert relerr(out["loss"], gd.out["loss"]) < 2e-6
    for k in TERMS:
        for g in (0, 1):
            assert relerr(out[k][g], gd.out[f"{k}{g}"]) < 5e-6, (k, g)
    if gd.mode == "label":
        for g in (0, 1):  # integer pairing contract: bit-exact
            assert np.array_equal(out["partners"][g], gd.out[f"partner{g}"])
    if gd.mode == "paired":
        for g in (0, 1):
            assert np.array_equal(out["partners"][g], gd.out[f"partner{g}"])
    if gd.training:
        worst, where = grad_errors(grads, gd.grads)
        assert worst < 2e-4, (worst, where)
        for k, v in out["new_stats"].items():
            assert relerr(v, gd.after[k]) < 5e-6, k


@pytest.mark.parametrize("name", ["label_tiny", "paired_tiny", "cluster_tiny"])
def test_oracle_float64_agrees(name):
    """the f32 oracle and an f64 run of the same restatement agree far inside the 1e-4 gate,
    i.e. the gate is not sitting on f32 rounding noise."""
    gd = Golden(name)
    o32, g32, _ = run_oracle(gd, torch.float32)
    o64, g64, _ = run_oracle(gd, torch.float64)
    assert relerr(o32["loss"], o64["loss"]) < 5e-6
    for k in ("rec", "kl_private", "kl_poe"):
        for g in (0, 1):
            assert relerr(o32[k][g], o64[k][g]) < 1e-5


def test_label_partner_rules():
    # rank-matched within label; PAD when the other group has fewer; ABSENT when label missing
    a = np.array([2, 0, 2, 1, 2, 5])
    b = np.array([0, 2, 3, 2, 0, 0])
    assert rs.label_partners(a, b).tolist() == [1, 0, 3, rs.PARTNER_ABSENT, rs.PARTNER_PAD, rs.PARTNER_ABSENT]
    assert rs.label_partners(b, a).tolist() == [1, 0, rs.PARTNER_ABSENT, 2, rs.PARTNER_PAD, rs.PARTNER_PAD]


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(50, dtype=torch.float64)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, eps=0.01, weight_decay=1e-6)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for t in range(1, 4):
        g = torch.randn(50, dtype=torch.float64)
        ref.grad = g.clone(); opt.step()
        p, m, v = rs.adam_step(p, g, m, v, t)
    assert relerr(p, ref.detach()) < 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("mode,n_batch", [("label", 0), ("paired", 0), ("cluster", 0), ("label", 4), ("cluster", 2), ("paired", 1)])
def test_oracle_matches_live_reference(mode, n_batch):
    from oracle import make_golden as mg, ref_harness as rh
    B, G, H, S, P, nl, drop, N = 40, (80, 64), 32, 25, 10, 6, 0.15, 90
    x_own, idx, labels, plan, eps_p, eps_q, masks = mg.synth(777, mode, B, G, H, S, P, nl, drop, N)
    xfull = [torch.cat([x_own[0], torch.zeros(B, G[1])], 1), torch.cat([torch.zeros(B, G[0]), x_own[1]], 1)]
    bcodes = [np.random.RandomState(5 + g).randint(0, max(n_batch, 1), B) for g in (0, 1)] if n_batch else None
    m = rh.build_reference(G, mode=mode, n_hidden=H, n_shared=S, n_private=P, dropout_rate=drop, plan=plan, n_labels=nl, seed=3,
                           n_batch=n_batch)
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    dm = {k: v.float() / (1 - drop) for k, v in masks.items()}
    batch = rh.make_batch(xfull, idx, labels=labels if mode == "label" else None, clabels=labels if mode == "cluster" else None,
                          batch=bcodes)
    ref = rh.run_reference(m, batch, eps_private=eps_p, eps_poe=eps_q, drop_masks=dm, kl_weight=0.5)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd0.items()}
    sub = rs.sub_plan(plan, idx[0], idx[1]) if mode != "label" else None
    out = rs.step(sd, x_own, mode=mode, n_shared=S, n_private=P, eps_private=eps_p, eps_poe=eps_q, labels=labels,
                  sub=sub, drop_masks=dm, kl_weight=0.5, batch_index=bcodes, n_batch=n_batch)
    out["loss"].backward()
    assert relerr(out["loss"], ref["loss"]) < 2e-6
    for k in TERMS:
        for g in (0, 1):
            assert relerr(out[k][g].reshape(ref[k][g].shape), ref[k][g]) < 5e-6, (k, g)
    worst, where = grad_errors({k: sd[k].grad for k in rs.param_names(sd)}, ref["grads"])
    assert worst < 2e-4, (worst, where)
    # output-dict key ORDER is part of the boundary (reference model/spvipes.py:539-551 unpacks positionally)
    assert ref["poe_keys"][0] == ["logtheta_loc", "logtheta_logvar", "logtheta_scale", "logtheta_qz", "logtheta_log_z", "logtheta_theta"]
    assert ref["private_keys"][0] == ["logtheta_loc", "logtheta_logvar", "logtheta_scale", "log_z", "theta", "qz"]
