"""TEST INFRASTRUCTURE ONLY — drive the UNMODIFIED reference module with injected noise.

The reference draws its reparameterisation noise from the global RNG through
torch.distributions.Normal.rsample (reference nn/networks.py:127, module/spVIPESmodule.py:
277,360,365,568,715).  To compare it with the oracle and the CUDA path on identical noise,
rsample is temporarily replaced by `loc + scale * eps` with eps chosen by CALLER:
Encoder.forward -> the per-encoder eps (the shared encoders' draw is discarded by the
reference), _poe2 -> discarded draw, the three *_poe functions -> the final PoE eps.
Nothing in the reference tree is modified.
"""
from __future__ import annotations

import contextlib
import sys
from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch.distributions import Normal

from . import ref_loader


def make_batch(x_full: Sequence[torch.Tensor], indices, labels=None, clabels=None, batch=None):
    """tuple-of-dicts minibatch in the layout scvi 0.20's AnnTorchDataset yields (all f32,
    [B,1] code columns) — reference module/spVIPESmodule.py:381-405."""
    out = []
    for g in (0, 1):
        B = x_full[g].shape[0]
        d = {
            "X": x_full[g].float(),
            "batch": torch.zeros(B, 1) if batch is None else torch.as_tensor(np.asarray(batch[g]).reshape(-1, 1), dtype=torch.float32),
            "groups": torch.full((B, 1), float(g)),
            "indices": torch.as_tensor(np.asarray(indices[g]).reshape(-1, 1), dtype=torch.float32),
        }
        if labels is not None:
            d["labels"] = torch.as_tensor(np.asarray(labels[g]).reshape(-1, 1), dtype=torch.float32)
        if clabels is not None:
            d["processed_transport_labels"] = torch.as_tensor(np.asarray(clabels[g]).reshape(-1, 1), dtype=torch.float32)
        out.append(d)
    return tuple(out)


def build_reference(genes, *, mode, n_hidden, n_shared, n_private, dropout_rate, plan=None, n_labels=None, seed=0, n_batch=0):
    cls, _ = ref_loader.load()
    G0, G1 = genes
    torch.manual_seed(seed)
    m = cls(
        groups_lengths={0: G0, 1: G1},
        groups_obs_names=[None, None],
        groups_var_names={0: None, 1: None},
        groups_obs_indices=[None, None],
        groups_var_indices=[np.arange(G0), np.arange(G0, G0 + G1)],
        transport_plan=plan if mode in ("paired", "cluster") else None,
        pair_data=(mode == "paired"),
        use_labels=(mode == "label"),
        n_labels=n_labels,
        n_batch=n_batch,
        n_hidden=n_hidden,
        n_dimensions_shared=n_shared,
        n_dimensions_private=n_private,
        dropout_rate=dropout_rate,
    )
    return m


class _MaskDrop(torch.nn.Module):
    """stands in for nn.Dropout on an encoder instance: multiplies by a supplied mask that
    already holds 0 or 1/(1-p) (what F.dropout does with its own Bernoulli draw)."""

    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask if self.training else x


@contextlib.contextmanager
def injected_noise(eps_private, eps_poe):
    calls = {"enc": 0, "poe": 0}
    orig = Normal.rsample

    def rsample(self, sample_shape=torch.Size()):
        caller = sys._getframe(1).f_code.co_name
        if caller == "forward":  # Encoder.forward: order private_0, shared_0, private_1, shared_1
            i = calls["enc"]
            calls["enc"] += 1
            if i % 2 == 0:
                return self.loc + self.scale * eps_private[i // 2].to(self.loc.dtype)
            return self.loc + self.scale * torch.zeros_like(self.loc)  # discarded by the reference
        if caller == "_poe2":
            return self.loc + self.scale * torch.zeros_like(self.loc)  # discarded by the reference
        if caller in ("_label_based_poe", "_paired_poe", "_cluster_based_poe"):
            i = calls["poe"]
            calls["poe"] += 1
            return self.loc + self.scale * eps_poe[i].to(self.loc.dtype)
        raise RuntimeError(f"unexpected rsample caller {caller}")

    Normal.rsample = rsample
    try:
        yield
    finally:
        Normal.rsample = orig


def run_reference(module, batch, *, eps_private, eps_poe, drop_masks: Optional[Dict] = None, kl_weight=1.0,
                  training=True, backward=True):
    """forward (+backward) of the unmodified reference; returns a dict shaped like
    oracle.restatement.step's plus grads by state_dict name."""
    if drop_masks:
        for (g, kind), mask in drop_masks.items():
            getattr(module, f"encoder_{g}_{kind}").drop = _MaskDrop(mask)
    module.train(training)
    module.zero_grad(set_to_none=True)
    with injected_noise(eps_private, eps_poe):
        inf, gen, lo = module(batch, loss_kwargs={"kl_weight": kl_weight})
    out = {
        "loss": lo.loss.detach(),
        "rec": [v.detach() for v in lo.reconstruction_loss.values()],
        "kl_private": [lo.kl_local["kl_divergence_groups_1_private"].detach(), lo.kl_local["kl_divergence_groups_2_private"].detach()],
        "kl_poe": [lo.kl_local["kl_divergence_groups_1_poe"].detach(), lo.kl_local["kl_divergence_groups_2_poe"].detach()],
        "library": [inf["library"][g].detach() for g in (0, 1)],
        "private_loc": [inf["private_stats"][g]["logtheta_loc"].detach() for g in (0, 1)],
        "private_logvar": [inf["private_stats"][g]["logtheta_logvar"].detach() for g in (0, 1)],
        "private_log_z": [inf["private_stats"][g]["log_z"].detach() for g in (0, 1)],
        "shared_loc": [inf["shared_stats"][g]["logtheta_loc"].detach() for g in (0, 1)],
        "shared_logvar": [inf["shared_stats"][g]["logtheta_logvar"].detach() for g in (0, 1)],
        "poe_loc": [inf["poe_stats"][g]["logtheta_loc"].detach() for g in (0, 1)],
        "poe_logvar": [inf["poe_stats"][g]["logtheta_logvar"].detach() for g in (0, 1)],
        "poe_scale": [inf["poe_stats"][g]["logtheta_scale"].detach() for g in (0, 1)],
        "poe_log_z": [inf["poe_stats"][g]["logtheta_log_z"].detach() for g in (0, 1)],
        "poe_keys": [list(inf["poe_stats"][g].keys()) for g in (0, 1)],
        "private_keys": [list(inf["private_stats"][g].keys()) for g in (0, 1)],
    }
    if backward:
        lo.loss.backward()
        out["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in module.named_parameters()}
    out["state_after"] = {k: v.detach().clone() for k, v in module.state_dict().items()}
    return out
