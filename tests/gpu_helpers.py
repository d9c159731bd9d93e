"""helpers for the -m gpu tests: drive spvipes_b200.engine.StepEngine from a golden fixture / synthetic case."""
import numpy as np
import torch


def engine_from_golden(gd, device="cuda", precision="fp32"):
    from spvipes_b200.engine import GroupBatch, Noise, StepEngine

    plan = gd.plan.to(device) if gd.mode != "label" else None
    eng = StepEngine((gd.G0, gd.G1), gd.H, gd.S, gd.P, gd.dropout, gd.mode, device, plan=plan, precision=precision,
                     n_batch=gd.n_batch)
    eng.load_state_dict(gd.sd)
    eng.set_kl_weight(gd.kl_weight)
    batches = []
    for g in (0, 1):
        X = torch.from_numpy(gd.x[g].numpy().astype(np.uint16)).to(device)
        labels = torch.from_numpy(gd.labels[g].astype(np.int32)).to(device) if gd.mode in ("label", "cluster") else None
        idx = torch.from_numpy(gd.idx[g].astype(np.int32)).to(device)
        bc = torch.from_numpy(np.asarray(gd.batch[g]).reshape(-1).astype(np.int32)).to(device) if gd.n_batch > 1 else None
        batches.append(GroupBatch(X=X, labels=labels, idx=idx, batch=bc))
    dm = gd.drop_masks()
    drop = None
    if dm is not None:
        drop = [torch.cat([dm[(g, "private")], dm[(g, "shared")]], dim=1).contiguous().to(device) for g in (0, 1)]
    noise = Noise(eps_private=[e.to(device) for e in gd.eps_private], eps_poe=[e.to(device) for e in gd.eps_poe], drop=drop)
    return eng, batches, noise


def engine_outputs(eng, ws):
    """same keys as oracle.restatement.step's output"""
    d = eng.d
    P, S = d.n_private, d.n_shared
    out = {"loss": eng.loss_out[0].cpu()}
    for k in ("library", "private_loc", "private_logvar", "private_log_z", "shared_loc", "shared_logvar", "poe_loc",
              "poe_logvar", "poe_scale", "poe_log_z", "rec", "kl_private", "kl_poe", "partners"):
        out[k] = []
    for w in ws:
        st = w.stats.cpu()
        out["library"].append(w.lib.cpu().unsqueeze(1))
        out["private_loc"].append(st[:, :P]); out["private_logvar"].append(st[:, P:2 * P])
        out["shared_loc"].append(st[:, 2 * P:2 * P + S]); out["shared_logvar"].append(st[:, 2 * P + S:])
        out["private_log_z"].append(w.zpriv.cpu()); out["poe_loc"].append(w.poe_loc.cpu())
        out["poe_logvar"].append(w.poe_lv.cpu()); out["poe_scale"].append(w.poe_scale.cpu())
        out["poe_log_z"].append(w.zpoe.cpu()); out["rec"].append(w.rec.cpu())
        out["kl_private"].append(w.klp.cpu()); out["kl_poe"].append(w.klq.cpu())
        out["partners"].append(w.partner.cpu().numpy().astype(np.int64))
    return out


def gate_consistent_grads(eng, ws, probe, grads, rerun, drop_masks=None):
    """oracle gradients to compare the engine's with.  `grads` / `probe` come from the oracle's plain pass; if the engine took the
    other decision on ReLU units whose pre-activation is zero to within tests.helpers.GATE_AMBIGUITY (and only on those:
    asserted), `rerun(gates)` evaluates the oracle again with those units' gates fixed to the engine's and returns its
    gradients.  Returns (grads, number of switched units)."""
    from tests.helpers import engine_gates, gate_consistent
    gates, switched = gate_consistent(probe, engine_gates(eng, ws, eng.d.n_hidden), drop_masks)
    if switched == 0:
        return grads, 0
    return rerun(gates), switched
