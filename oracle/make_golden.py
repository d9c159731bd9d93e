"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden
Each fixture holds the inputs (own-gene counts, within-group indices, labels, transport
plan, injected noise, dropout keep-masks, initial state_dict) and what the reference
module (src/spVIPES/module/spVIPESmodule.py via oracle/scvi_stub) produced for them: loss,
the 2 reconstruction + 4 KL terms, library, latent statistics, pairing indices, parameter
gradients and the BatchNorm running statistics after the step.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as rh  # noqa: E402
from oracle import restatement as rs  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
# fixtures of SURVEY.md section 8(f) rows the CUDA path does not cover yet: kept out of tests/golden/, whose files the GPU
# parity tests enumerate
OUT_NEXT = os.path.join(os.path.dirname(HERE), "tests", "golden_next")

CASES = {
    # name: (mode, B, (G0, G1), H, S, P, n_labels, dropout, N_per_group, training)
    "label_tiny": ("label", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, True),
    "paired_tiny": ("paired", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, True),
    "cluster_tiny": ("cluster", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, True),
    "label_tutorial_dims": ("label", 32, (70, 45), 48, 10, 7, 5, 0.0, 64, True),
    "paired_tutorial_dims": ("paired", 32, (70, 45), 48, 10, 7, 5, 0.2, 64, True),
    "label_medium": ("label", 96, (300, 350), 64, 25, 10, 10, 0.1, 400, True),
    "label_eval": ("label", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, False),
    "cluster_eval": ("cluster", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, False),
}


# batch covariates (n_batch > 1): name -> (spec as in CASES, n_batch)
CASES_NEXT = {
    "label_batch3_tiny": (("label", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, True), 3),
    "paired_batch2_eval": (("paired", 24, (50, 60), 32, 25, 10, 4, 0.1, 40, False), 2),
}


def synth(case_seed, mode, B, G, H, S, P, nl, drop, N):
    g = torch.Generator().manual_seed(case_seed)
    rate = [torch.rand(1, G[i], generator=g) ** 3 * 6.0 for i in (0, 1)]
    x_own = [torch.poisson(rate[i].expand(B, -1) * (0.5 + torch.rand(B, 1, generator=g)), generator=g) for i in (0, 1)]
    idx = [torch.randperm(N, generator=g)[:B].numpy().astype(np.int64) for _ in (0, 1)]
    labels = [torch.randint(0, nl, (B,), generator=g).numpy().astype(np.int64) for _ in (0, 1)]
    if nl >= 4:
        labels[0][labels[0] == nl - 1] = nl - 2  # last label absent from group 0 -> "absent partner" branch
        labels[1][labels[1] == 0] = 1  # label 0 absent from group 1
    plan = torch.rand(N, N, generator=g)
    plan[plan < 0.3] = 0.0
    eps_p = [torch.randn(B, P, generator=g) for _ in (0, 1)]
    eps_q = [torch.randn(B, S, generator=g) for _ in (0, 1)]
    keep = 1.0 - drop
    masks = {(gg, k): (torch.rand(B, H, generator=g) < keep) for gg in (0, 1) for k in ("private", "shared")}
    return x_own, idx, labels, plan, eps_p, eps_q, masks


def make(name, spec, seed, n_batch=0, out_dir=None):
    mode, B, G, H, S, P, nl, drop, N, training = spec
    x_own, idx, labels, plan, eps_p, eps_q, masks = synth(seed, mode, B, G, H, S, P, nl, drop, N)
    xfull = [torch.cat([x_own[0], torch.zeros(B, G[1])], 1), torch.cat([torch.zeros(B, G[0]), x_own[1]], 1)]
    bcodes = None
    if n_batch > 1:
        gb = torch.Generator().manual_seed(seed + 5000)
        bcodes = [torch.randint(0, n_batch, (B,), generator=gb).numpy().astype(np.int64) for _ in (0, 1)]
    m = rh.build_reference(G, mode=mode, n_hidden=H, n_shared=S, n_private=P, dropout_rate=drop, plan=plan, n_labels=nl, seed=seed,
                           n_batch=n_batch)
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():  # move BN affine / running stats off their defaults so that they matter
        for k, p in m.named_parameters():
            if k.endswith(".1.weight") or k.endswith(".1.bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
        for k, b in m.named_buffers():
            if k.endswith("running_mean"):
                b.add_(0.1 * torch.randn(b.shape, generator=g))
            if k.endswith("running_var"):
                b.mul_(1.0 + 0.3 * torch.rand(b.shape, generator=g))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    keep = 1.0 - drop
    dm = {k: v.float() / keep for k, v in masks.items()}
    batch = rh.make_batch(xfull, idx, labels=labels if mode == "label" else None, clabels=labels if mode == "cluster" else None,
                          batch=bcodes)
    kl_w = 0.37
    ref = rh.run_reference(m, batch, eps_private=eps_p, eps_poe=eps_q, drop_masks=dm if drop > 0 else None,
                           kl_weight=kl_w, training=training, backward=training)
    blob = {
        "meta_mode": np.array(mode), "meta_dims": np.array([B, G[0], G[1], H, S, P, nl, N], dtype=np.int64),
        "meta_dropout": np.array(drop), "meta_kl_weight": np.array(kl_w), "meta_training": np.array(training),
        "plan": plan.numpy(), "meta_n_batch": np.array(n_batch, dtype=np.int64),
    }
    for gi in (0, 1):
        if bcodes is not None:
            blob[f"batch{gi}"] = bcodes[gi]
        blob[f"x{gi}"] = x_own[gi].numpy().astype(np.uint16)
        blob[f"idx{gi}"] = idx[gi]
        blob[f"labels{gi}"] = labels[gi]
        blob[f"eps_private{gi}"] = eps_p[gi].numpy()
        blob[f"eps_poe{gi}"] = eps_q[gi].numpy()
        for k in ("private", "shared"):
            blob[f"keep_{gi}_{k}"] = masks[(gi, k)].numpy()
        for k in ("rec", "kl_private", "kl_poe", "library", "private_loc", "private_logvar", "private_log_z",
                  "shared_loc", "shared_logvar", "poe_loc", "poe_logvar", "poe_scale", "poe_log_z"):
            blob[f"out_{k}{gi}"] = ref[k][gi].numpy()
    blob["out_loss"] = ref["loss"].numpy()
    # integer pairing contract
    if mode == "label":
        blob["out_partner0"] = rs.label_partners(labels[0], labels[1])
        blob["out_partner1"] = rs.label_partners(labels[1], labels[0])
    else:
        sub = rs.sub_plan(plan, idx[0], idx[1])
        blob["out_partner0"] = torch.argmax(sub, dim=1).numpy()  # reference :526
        blob["out_partner1"] = torch.argmax(sub, dim=0).numpy()  # reference :527
    for k, v in sd0.items():
        blob["sd/" + k] = v.numpy()
    if training:
        for k, v in ref["grads"].items():
            blob["grad/" + k] = v.numpy()
    for k, v in ref["state_after"].items():
        if "running" in k:
            blob["after/" + k] = v.numpy()
    out_dir = out_dir or OUT
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: loss={float(ref['loss']):.6f}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    if "--next-only" not in sys.argv:
        for i, (name, spec) in enumerate(CASES.items()):
            make(name, spec, 100 + i)
    for i, (name, (spec, nb)) in enumerate(CASES_NEXT.items()):
        make(name, spec, 300 + i, n_batch=nb, out_dir=OUT_NEXT)
