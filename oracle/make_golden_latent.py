"""TEST INFRASTRUCTURE ONLY - fixtures for the model-level latent extraction.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden_latent

The reference's model class (src/spVIPES/model/spvipes.py) cannot be imported here (scvi-tools / anndata absent), so its batching is
restated below, line by line, around the UNMODIFIED reference module (module/spVIPESmodule.py through oracle/scvi_stub, eval mode):
  * get_latent_representation            model/spvipes.py:424-525  (drop_last False; the paired OT mode takes the cycling path)
  * _process_batches                     :537-576   (positional unpack of the stats dicts, the SAMPLED log_z are collected)
  * _process_all_cells_with_cycling      :578-626
  * _format_results                      :628-650   (truncation, argsort of group 2 by obs["indices"])
  * ConcatDataLoader(shuffle=False)      dataloaders/_concat_dataloader.py:101-110 (zip(largest, cycle(other)))
with the reparameterisation noise injected per minibatch (generator seeded 1000 + batch number; oracle/ref_harness.py).
"""
from __future__ import annotations

import os
import sys
from itertools import cycle

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden_latent")

CASES = {
    # name: (mode, (n0, n1), (G0, G1), H, S, P, n_labels, batch_size)
    "latent_label_ragged": ("label", (70, 53), (40, 36), 32, 12, 6, 4, 32),
    "latent_paired_cycling": ("paired", (50, 34), (40, 36), 32, 12, 6, 4, 16),
    "latent_cluster_equal": ("cluster", (48, 48), (40, 36), 32, 12, 6, 3, 16),
}


def batch_noise(k, B0, B1, P, S):
    g = torch.Generator().manual_seed(1000 + k)
    return [torch.randn(B, P, generator=g) for B in (B0, B1)], [torch.randn(B, S, generator=g) for B in (B0, B1)]


def generate(name):
    mode, n, G, H, S, P, nl, bs = CASES[name]
    g = torch.Generator().manual_seed(sum(name.encode()) % 1000 + 5)  # stable across processes (hash() is salted)
    rate = [torch.rand(1, G[i], generator=g) ** 3 * 6.0 for i in (0, 1)]
    x = [torch.poisson(rate[i].expand(n[i], -1) * (0.5 + torch.rand(n[i], 1, generator=g)), generator=g) for i in (0, 1)]
    labels = [torch.randint(0, nl, (n[i],), generator=g).numpy().astype(np.int64) for i in (0, 1)]
    plan = torch.rand(n[0], n[1], generator=g)
    plan[plan < 0.3] = 0.0
    m = rh.build_reference(G, mode=mode, n_hidden=H, n_shared=S, n_private=P, dropout_rate=0.1, plan=plan, n_labels=nl, seed=11)
    # non-trivial BatchNorm running statistics (eval mode uses them)
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(0.3 * torch.randn(v.shape, generator=g))
            elif k.endswith("running_var"):
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
    m.eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    # rows of the combined matrix: group 0 first (prepare_adatas), obs["indices"] = within-group position
    gil = [list(range(n[0])), list(range(n[0], n[0] + n[1]))]
    within = np.concatenate([np.arange(n[0]), np.arange(n[1])])
    grp = np.concatenate([np.zeros(n[0], int), np.ones(n[1], int)])
    res = {k: [] for k in ("s1", "s2", "p1", "p2", "i1", "i2")}
    counter = [0]

    def process(index_lists):  # _process_batches over ConcatDataLoader(shuffle=False, drop_last=False)
        chunks = [[np.asarray(gi)[k:k + bs] for k in range(0, len(gi), bs)] for gi in index_lists]
        largest = int(np.argmax([len(c) for c in chunks]))
        its = [iter(c) if gidx == largest else cycle(c) for gidx, c in enumerate(chunks)]
        for st in zip(*its):
            loc = [within[st[i]] for i in (0, 1)]
            assert all((grp[st[i]] == i).all() for i in (0, 1))
            xf = [torch.cat([x[0][loc[0]], torch.zeros(len(loc[0]), G[1])], 1), torch.cat([torch.zeros(len(loc[1]), G[0]), x[1][loc[1]]], 1)]
            batch = rh.make_batch(xf, loc, labels=[labels[i][loc[i]] for i in (0, 1)] if mode == "label" else None,
                                  clabels=[labels[i][loc[i]] for i in (0, 1)] if mode == "cluster" else None)
            eps_p, eps_q = batch_noise(counter[0], len(loc[0]), len(loc[1]), P, S)
            counter[0] += 1
            with torch.no_grad(), rh.injected_noise(eps_p, eps_q):
                inp = m._get_inference_input(batch)
                out = m.inference(**inp)
            _, _, _, _, z1, _ = out["poe_stats"][0].values()       # :539
            _, _, _, _, z2, _ = out["poe_stats"][1].values()       # :540
            _, _, _, pz1, _, _ = out["private_stats"][0].values()  # :546
            _, _, _, pz2, _, _ = out["private_stats"][1].values()
            res["s1"].append(z1.cpu()); res["s2"].append(z2.cpu()); res["p1"].append(pz1.cpu()); res["p2"].append(pz2.cpu())
            res["i1"].append(batch[0]["indices"].cpu()); res["i2"].append(batch[1]["indices"].cpu())

    if mode == "paired":  # _process_all_cells_with_cycling :578-626
        lo, hi = min(n), max(n)
        for start in range(0, hi, lo):
            process([[gil[i][(start + j) % n[i]] for j in range(lo)] for i in (0, 1)])
    else:
        process(gil)
    i2 = torch.cat(res["i2"]).numpy().flatten()[:n[1]]  # _format_results :628-650
    p = [torch.cat(res["p1"]).numpy()[:n[0]], torch.cat(res["p2"]).numpy()[:n[1]]]
    sh = [torch.cat(res["s1"]).numpy()[:n[0]], torch.cat(res["s2"]).numpy()[:n[1]]]
    order = np.argsort(i2)
    out = {"meta_mode": mode, "meta_dims": np.array([n[0], n[1], G[0], G[1], H, S, P, nl, bs]), "plan": plan.numpy(),
           "x0": x[0].numpy().astype(np.int32), "x1": x[1].numpy().astype(np.int32), "labels0": labels[0], "labels1": labels[1],
           "n_batches": np.array(counter[0]),
           "shared0": sh[0], "shared1": sh[1], "private0": p[0], "private1": p[1], "shared1_reordered": sh[1][order],
           "private1_reordered": p[1][order]}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "batches", counter[0], {k: v.shape for k, v in out.items() if k.startswith(("shared", "private"))})


if __name__ == "__main__":
    for name in CASES:
        generate(name)
