"""-m gpu: the drop-in module / model API (spvipes_b200.module.spVIPESmodule, spvipes_b200.model.spVIPES) driven the way
scvi's TrainingPlan and the reference's model class drive the reference module."""
import numpy as np
import pandas as pd
import pytest
import torch

from tests.helpers import Golden, grad_errors, relerr

pytestmark = pytest.mark.gpu


def _module_from_golden(gd):
    from spvipes_b200.engine import Noise
    from spvipes_b200.module import spVIPESmodule

    G0, G1 = gd.G0, gd.G1
    m = spVIPESmodule(groups_lengths={0: G0, 1: G1}, groups_obs_names=[None, None], groups_var_names={0: None, 1: None},
                      groups_obs_indices=[None, None], groups_var_indices=[np.arange(G0), np.arange(G0, G0 + G1)],
                      transport_plan=gd.plan if gd.mode != "label" else None, pair_data=gd.mode == "paired",
                      use_labels=gd.mode == "label", n_labels=gd.n_labels, n_hidden=gd.H, n_dimensions_shared=gd.S,
                      n_dimensions_private=gd.P, dropout_rate=gd.dropout)
    missing, unexpected = m.load_state_dict(gd.sd, strict=True), None
    dm = gd.drop_masks()
    drop = None
    if dm is not None:
        drop = [torch.cat([dm[(g, "private")], dm[(g, "shared")]], dim=1).contiguous().cuda() for g in (0, 1)]
    m._noise = Noise([e.cuda() for e in gd.eps_private], [e.cuda() for e in gd.eps_poe], drop)
    B = gd.B
    xfull = [torch.cat([gd.x[0].float(), torch.zeros(B, G1)], 1), torch.cat([torch.zeros(B, G0), gd.x[1].float()], 1)]
    batch = []
    for g in (0, 1):  # the layout scvi 0.20's AnnTorchDataset yields (all float32, [B, 1] code columns)
        d = {"X": xfull[g], "batch": torch.zeros(B, 1), "groups": torch.full((B, 1), float(g)),
             "indices": torch.as_tensor(gd.idx[g].reshape(-1, 1), dtype=torch.float32)}
        if gd.mode == "label":
            d["labels"] = torch.as_tensor(gd.labels[g].reshape(-1, 1), dtype=torch.float32)
        if gd.mode == "cluster":
            d["processed_transport_labels"] = torch.as_tensor(gd.labels[g].reshape(-1, 1), dtype=torch.float32)
        batch.append(d)
    return m, tuple(batch)


@pytest.mark.parametrize("name", ["label_tiny", "paired_tiny", "cluster_tiny"])
def test_module_forward_backward_like_training_plan(name):
    gd = Golden(name)
    m, batch = _module_from_golden(gd)
    assert sorted(m.state_dict().keys()) == sorted(gd.sd.keys())  # reference checkpoint names, 1:1
    m.train()
    inf, gen, lo = m(batch, loss_kwargs={"kl_weight": gd.kl_weight})
    assert relerr(lo.loss.detach().cpu(), gd.out["loss"]) < 1e-4
    assert list(inf["poe_stats"][0].keys()) == ["logtheta_loc", "logtheta_logvar", "logtheta_scale", "logtheta_qz",
                                                "logtheta_log_z", "logtheta_theta"]
    assert list(inf["private_stats"][0].keys()) == ["logtheta_loc", "logtheta_logvar", "logtheta_scale", "log_z", "theta", "qz"]
    for g, k in enumerate(("reconst_loss_groups_1_poe", "reconst_loss_groups_2_poe")):
        assert relerr(lo.reconstruction_loss[k].cpu(), gd.out[f"rec{g}"]) < 1e-4
    assert relerr(lo.kl_local["kl_divergence_groups_2_poe"].cpu(), gd.out["kl_poe1"]) < 1e-4
    lo.loss.backward()
    got = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    worst, where = grad_errors(got, gd.grads)
    assert worst < 2e-3, (worst, where)
    # a torch optimiser updates the engine's flat store through the parameter views
    before = m.engine.params.flat.clone()
    torch.optim.Adam(m.parameters(), lr=1e-3, eps=0.01, weight_decay=1e-6).step()
    assert not torch.equal(before, m.engine.params.flat)


def test_model_api_train_and_latent():
    from spvipes_b200 import synth
    from spvipes_b200.model import GroupedData, prepare_adatas, spVIPES

    n, G = (700, 600), (96, 80)
    data = synth.make_counts(n, G, n_labels=4, device="cuda", seed=5)
    ads = {}
    for gi, key in enumerate(("mouse", "human")):
        X = data.X[gi].cpu().numpy().astype(np.float32)
        obs = pd.DataFrame({"cell_type": [f"t{int(v)}" for v in data.labels[gi].cpu().numpy()]})
        ads[key] = GroupedData(X=X, obs=obs, var_names=[f"g{j}" for j in range(G[gi])])
    adata = prepare_adatas(ads)
    assert adata.X.shape == (sum(n), sum(G)) and list(adata.uns["groups_lengths"].values()) == list(G)
    spVIPES.setup_anndata(adata, groups_key="groups", label_key="cell_type")
    model = spVIPES(adata, n_hidden=64, n_dimensions_shared=12, n_dimensions_private=6, dropout_rate=0.1)
    gil = [list(ix) for ix in adata.uns["groups_obs_indices"]]
    model.train(gil, max_epochs=12, batch_size=128, train_size=0.9, n_epochs_kl_warmup=None)  # constant KL weight 1
    h = model.history["train_loss_epoch"]
    assert len(h) == 12 and np.isfinite(h).all() and h[-1] < h[0], h
    lat = model.get_latent_representation(gil, batch_size=256)
    assert lat["shared"][0].shape == (n[0], 12) and lat["shared"][1].shape == (n[1], 12)
    assert lat["private"][0].shape == (n[0], 6) and lat["private_reordered"][1].shape == (n[1], 6)
    assert np.isfinite(lat["shared"][0]).all() and np.isfinite(lat["private"][1]).all()
    with pytest.raises(RuntimeError):  # as the reference: torch.cat of the empty shared list (model/spvipes.py:542-544, 634)
        model.get_latent_representation(gil, batch_size=256, normalized=True)
    load = model.get_loadings()
    assert load[(0, "shared")].shape == (G[0], 12) and load[(1, "private")].shape == (G[1], 6)
