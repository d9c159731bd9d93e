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


@pytest.mark.parametrize("mode,precision", [("label", "bf16"), ("paired", "fp32")])
def test_plugin_call_captured_equals_eager(mode, precision):
    """`module(batch, loss_kwargs)` -> `loss.backward()` -> torch Adam, driven like scvi's TrainingPlan with pinned host
    minibatches in scvi's layout (X float32 [B, G0 + G1], [B, 1] float code columns): the captured fast path (two CUDA graphs
    per buffer set, strided column copy, lazily built output dicts) leaves the same parameters as the eager path."""
    from spvipes_b200 import synth
    from spvipes_b200.module import spVIPESmodule
    B, G0, G1, H, NL, STEPS = 128, 300, 260, 64, 5, 7
    data = synth.make_counts((4 * B, 4 * B), (G0, G1), NL, device="cuda", seed=9)
    plan = synth.make_plan(4 * B, 4 * B, data.labels[0], data.labels[1], NL, device="cuda").cpu() if mode == "paired" else None
    gen = torch.Generator().manual_seed(3)
    steps = []
    for s in range(STEPS):
        batch = []
        for g in (0, 1):
            r = torch.randperm(4 * B, generator=gen)[:B]
            X = torch.zeros(B, G0 + G1)
            X[:, (0 if g == 0 else G0):(G0 if g == 0 else G0 + G1)] = data.X[g].cpu().to(torch.int32)[r].float()
            d = {"X": X.pin_memory(), "batch": torch.zeros(B, 1), "groups": torch.full((B, 1), float(g)),
                 "indices": r.float().reshape(-1, 1).pin_memory()}
            if mode == "label":
                d["labels"] = data.labels[g].cpu()[r].float().reshape(-1, 1).pin_memory()
            batch.append(d)
        steps.append(tuple(batch))
    res = []
    for captured in (False, True):
        torch.manual_seed(1234)
        m = spVIPESmodule(groups_lengths={0: G0, 1: G1}, groups_obs_names=[None, None], groups_var_names={0: None, 1: None},
                          groups_obs_indices=[None, None], groups_var_indices=[np.arange(G0), np.arange(G0, G0 + G1)],
                          transport_plan=plan, pair_data=mode == "paired", use_labels=mode == "label", n_labels=NL, n_hidden=H,
                          dropout_rate=0.1, precision=precision)
        m.capture_steps = captured
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, eps=0.01, weight_decay=1e-6)
        losses = []
        for s in range(STEPS):
            opt.zero_grad()
            inf, gen_out, lo = m(steps[s], loss_kwargs={"kl_weight": 0.1 * s})
            lo.loss.backward()
            opt.step()
            losses.append(float(lo.loss))
        torch.cuda.synchronize()
        assert list(inf["private_stats"][0].keys()) == ["logtheta_loc", "logtheta_logvar", "logtheta_scale", "log_z", "theta", "qz"]
        assert tuple(inf["poe_stats"][1]["logtheta_theta"].shape) == (B, 25)  # lazily built entries materialise on access
        res.append((m.engine.params.flat.clone(), m.engine.buffers.flat.clone(), losses))
    assert all(np.isfinite(res[1][2]))
    assert float((res[0][0] - res[1][0]).abs().max()) <= 1e-6 * float(res[0][0].abs().max()), "parameters differ"
    assert float((res[0][1] - res[1][1]).abs().max()) <= 1e-5 * float(res[0][1].abs().max()), "running statistics differ"
    assert max(abs(a - b) for a, b in zip(res[0][2], res[1][2])) <= 1e-5 * abs(res[0][2][0])


def test_flat_adam_equals_torch_adam():
    """spvipes_b200.optim.FlatAdam (one fused launch over the flat buffers, refreshes the 16-bit operand copies) walks the same
    trajectory as torch.optim.Adam with scvi's TrainingPlan hyper-parameters, driven through the plugin call"""
    from spvipes_b200.optim import FlatAdam
    gd = Golden("label_medium")
    res = []
    for flat in (False, True):
        torch.manual_seed(7)
        m, batch = _module_from_golden(gd)
        m._noise = None
        m.engine.dropout_rate = 0.0
        m.train()
        opt = FlatAdam(m) if flat else torch.optim.Adam(m.parameters(), lr=1e-3, eps=0.01, weight_decay=1e-6)
        for s in range(6):
            opt.zero_grad()
            _, _, lo = m(batch, loss_kwargs={"kl_weight": 0.3})
            lo.loss.backward()
            opt.step()
        torch.cuda.synchronize()
        res.append(m.engine.params.flat.clone())
    assert float((res[0] - res[1]).abs().max()) <= 2e-6 * float(res[0].abs().max())


@pytest.mark.parametrize("name", ["latent_label_ragged", "latent_paired_cycling", "latent_cluster_equal"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_get_latent_representation_matches_reference(name, precision):
    """model-level latent extraction (reference model/spvipes.py:424-650) against fixtures produced by the UNMODIFIED reference
    module driven through the reference's batching (oracle/make_golden_latent.py): sequential minibatches with the shorter loader
    cycled, ragged last batches, the paired mode's cycling over chunks of min(n) cells, truncation and the argsort of group 2,
    same injected noise per minibatch.  Values <= 1e-3 (north_star's latent gate), eval mode (running statistics)."""
    import os
    from spvipes_b200.engine import Noise
    from spvipes_b200.model import GroupedData, prepare_adatas, spVIPES
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_latent", name + ".npz"))
    mode = str(z["meta_mode"])
    n0, n1, G0, G1, H, S, P, nl, bs = (int(v) for v in z["meta_dims"])
    ads = {}
    for gi, key in enumerate(("a_first", "b_second")):
        obs = pd.DataFrame({"cell_type": [f"t{int(v)}" for v in z[f"labels{gi}"]]})
        if mode == "cluster":
            obs["processed_transport_labels"] = z[f"labels{gi}"]
        ads[key] = GroupedData(X=z[f"x{gi}"].astype(np.float32), obs=obs, var_names=[f"g{j}" for j in range((G0, G1)[gi])])
    adata = prepare_adatas(ads)
    if mode != "label":
        adata.uns["transport_plan"] = z["plan"]
    spVIPES.setup_anndata(adata, groups_key="groups", label_key="cell_type" if mode == "label" else None,
                          transport_plan_key="transport_plan" if mode != "label" else None, match_clusters=mode == "cluster")
    model = spVIPES(adata, n_hidden=H, n_dimensions_shared=S, n_dimensions_private=P, dropout_rate=0.1, precision=precision)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model.module.load_state_dict(sd, strict=True)

    def noise(k, B0, B1):  # the generator of oracle/make_golden_latent.batch_noise
        g = torch.Generator().manual_seed(1000 + k)
        ep = [torch.randn(B, P, generator=g) for B in (B0, B1)]
        eq = [torch.randn(B, S, generator=g) for B in (B0, B1)]
        return Noise([e.cuda() for e in ep], [e.cuda() for e in eq], None)

    gil = [list(ix) for ix in adata.uns["groups_obs_indices"]]
    lat = model.get_latent_representation(gil, batch_size=bs, _noise_fn=noise)
    for key, got in (("shared0", lat["shared"][0]), ("shared1", lat["shared"][1]), ("private0", lat["private"][0]),
                     ("private1", lat["private"][1]), ("shared1_reordered", lat["shared_reordered"][1]),
                     ("private1_reordered", lat["private_reordered"][1])):
        assert got.shape == z[key].shape, key
        assert relerr(got, z[key]) < 1e-3, (key, relerr(got, z[key]))


def test_model_api_with_batch_covariate_and_step_warmup():
    """setup_anndata(batch_key=...) with two batches: the one-hot batch code reaches the kernels through train() and
    get_latent_representation(); n_steps_kl_warmup drives the KL weight per step"""
    from spvipes_b200 import synth
    from spvipes_b200.model import GroupedData, prepare_adatas, spVIPES
    n, G = (500, 420), (64, 72)
    data = synth.make_counts(n, G, n_labels=3, device="cuda", seed=15)
    ads = {}
    rng = np.random.RandomState(0)
    for gi, key in enumerate(("mouse", "human")):
        obs = pd.DataFrame({"cell_type": [f"t{int(v)}" for v in data.labels[gi].cpu().numpy()], "lab": rng.choice(["b0", "b1"], n[gi])})
        ads[key] = GroupedData(X=data.X[gi].cpu().numpy().astype(np.float32), obs=obs, var_names=[f"g{j}" for j in range(G[gi])])
    adata = prepare_adatas(ads)
    spVIPES.setup_anndata(adata, groups_key="groups", label_key="cell_type", batch_key="lab")
    for precision in ("fp32", "bf16"):
        model = spVIPES(adata, n_hidden=64, n_dimensions_shared=12, n_dimensions_private=6, dropout_rate=0.1, precision=precision)
        assert model.module.engine.d.nb == 2
        assert model.module.state_dict()["encoder_0_private.fc1.weight"].shape == (64, G[0] + 2)
        gil = [list(ix) for ix in adata.uns["groups_obs_indices"]]
        model.train(gil, max_epochs=8, batch_size=128, train_size=0.9, n_epochs_kl_warmup=None)  # constant KL weight 1
        h = model.history["train_loss_epoch"]
        assert np.isfinite(h).all() and h[-1] < h[0], h
        model.train(gil, max_epochs=2, batch_size=128, train_size=0.9, n_steps_kl_warmup=10)
        # 3 steps per epoch (450 / 378 training rows, drop_last, the smaller group cycled): step k runs at weight k / 10
        assert abs(float(model.module.engine.kl_weight) - 0.5) < 1e-6
        model.train(gil, max_epochs=5, batch_size=128, train_size=0.9, n_steps_kl_warmup=10)
        assert float(model.module.engine.kl_weight) == 1.0  # the 10 warm-up steps are over
        assert np.isfinite(model.history["train_loss_epoch"]).all()
        lat = model.get_latent_representation(gil, batch_size=200)
        assert lat["shared"][1].shape == (n[1], 12) and np.isfinite(lat["private"][0]).all()
        assert model.get_loadings()[(0, "private")].shape == (G[0], 6)


def test_model_api_cluster_mode_derives_transport_labels():
    """setup_anndata(match_clusters=True, transport_plan_key=...) without ready-made labels: process_transport_plan (reference
    model/spvipes.py:26-162, :362-370) derives `processed_transport_labels` with the built-in clustering, and cluster-based PoE
    trains on them"""
    from spvipes_b200 import synth
    from spvipes_b200.model import GroupedData, prepare_adatas, spVIPES

    n, G = (300, 260), (64, 48)
    data = synth.make_counts(n, G, n_labels=3, device="cuda", seed=9)
    ads = {}
    for gi, key in enumerate(("a", "b")):
        ads[key] = GroupedData(X=data.X[gi].cpu().numpy().astype(np.float32), obs=pd.DataFrame({"dummy": np.zeros(n[gi])}),
                               var_names=[f"g{j}" for j in range(G[gi])])
    adata = prepare_adatas(ads)
    plan = synth.make_plan(n[0], n[1], data.labels[0].cpu(), data.labels[1].cpu(), 3, device="cpu", seed=3)
    adata.uns["transport_plan"] = plan.numpy()
    spVIPES.setup_anndata(adata, groups_key="groups", transport_plan_key="transport_plan", match_clusters=True)
    lab = adata.obs["processed_transport_labels"]
    assert lab.notna().all() and 2 <= len(lab.cat.categories) <= 40 and set(adata.uns["optimal_resolutions"]) == {"a", "b"}
    model = spVIPES(adata, n_hidden=32, n_dimensions_shared=8, n_dimensions_private=4, dropout_rate=0.1)
    assert model.module.use_transport_plan and not model.module.pair_data
    gil = [list(ix) for ix in adata.uns["groups_obs_indices"]]
    model.train(gil, max_epochs=3, batch_size=64, train_size=0.9, n_epochs_kl_warmup=None)
    assert np.isfinite(model.history["train_loss_epoch"]).all()
