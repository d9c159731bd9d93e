"""TEST INFRASTRUCTURE ONLY — CPU restatement of the spVIPES per-minibatch training step.

This is the ORACLE for the hot path named by BASELINE.json:north_star.  It is an
independent, functional (state_dict in, tensors out) restatement in plain PyTorch of

  * spVIPESmodule.inference      reference src/spVIPES/module/spVIPESmodule.py:425-472
  * Encoder.forward              reference src/spVIPES/nn/networks.py:85-140
  * _label_based_poe / _poe2     reference module/spVIPESmodule.py:583-718, 282-379
  * _paired_poe / _product_of_experts   reference :511-581, _get_batch_transport_plans :474-482
  * _cluster_based_poe           reference :184-280
  * generative                   reference :720-771
  * LinearDecoderSPVIPE.forward  reference nn/networks.py:264-335
  * loss                         reference :809-899
  * scvi-tools 0.20.0 (pyproject.toml:25, source absent): FCLayers (Linear ->
    BatchNorm1d(momentum=0.01, eps=0.001)), NegativeBinomialMixture.log_prob ->
    log_mixture_nb(eps=1e-8, shared theta), torch kl_divergence(Normal, Normal).

It takes the reparameterisation noise and the dropout masks as explicit inputs (the
reference draws them from the global RNG, including draws whose results it discards,
module/spVIPESmodule.py:359-366), so the reference, this oracle and the CUDA path can be
compared on identical noise.

Pinning: checked in tests/test_oracle_vs_reference.py against the UNMODIFIED reference
files run through oracle/scvi_stub (authoring container only) and against the committed
fixtures in tests/golden/ that were generated from them (oracle/make_golden.py).  The
scvi-tools arithmetic itself is restated from the published 0.20.0 behaviour and has no
reference test or golden vector behind it:  **parity unpinned** at that boundary.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this file.  The product package (spvipes_b200/) never does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

NB_EPS = 1e-8  # scvi log_mixture_nb eps
ENC_BN_EPS, ENC_BN_MOM = 1e-5, 0.1  # nn.BatchNorm1d defaults, reference nn/networks.py:76,82
DEC_BN_EPS, DEC_BN_MOM = 1e-3, 0.01  # scvi FCLayers BatchNorm1d(momentum=0.01, eps=0.001)


# ----------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------
def batch_norm(x, weight, bias, running_mean, running_var, training, eps, momentum, new_stats=None, key=None):
    """nn.BatchNorm1d forward.  Training: biased batch variance for normalisation, running
    stats updated with the UNBIASED variance (returned in new_stats, inputs not mutated)."""
    if training:
        n = x.shape[0]
        mean = x.mean(0)
        var = ((x - mean) ** 2).mean(0)
        if new_stats is not None:
            unb = var * (n / max(n - 1, 1))
            new_stats[key + ".running_mean"] = ((1 - momentum) * running_mean + momentum * mean).detach()
            new_stats[key + ".running_var"] = ((1 - momentum) * running_var + momentum * unb).detach()
    else:
        mean, var = running_mean, running_var
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def kl_std_normal(loc, scale):
    """torch.distributions.kl.kl_normal_normal(Normal(loc, scale), Normal(0, 1)).sum(1)."""
    var_ratio = scale ** 2
    t1 = loc ** 2
    return (0.5 * (var_ratio + t1 - 1 - torch.log(var_ratio))).sum(1)


def log_mixture_nb(x, mu_1, mu_2, theta, pi_logits, eps=NB_EPS):
    """scvi.distributions._negative_binomial.log_mixture_nb, shared-theta branch."""
    theta = theta.view(1, -1)
    l1 = torch.log(theta + mu_1 + eps)
    l2 = torch.log(theta + mu_2 + eps)
    lg = torch.lgamma(x + theta) - torch.lgamma(theta) - torch.lgamma(x + 1)
    log_nb_1 = theta * (torch.log(theta + eps) - l1) + x * (torch.log(mu_1 + eps) - l1) + lg
    log_nb_2 = theta * (torch.log(theta + eps) - l2) + x * (torch.log(mu_2 + eps) - l2) + lg
    lse = torch.logsumexp(torch.stack((log_nb_1, log_nb_2 - pi_logits)), dim=0)
    return lse - F.softplus(-pi_logits)


# ----------------------------------------------------------------------------------------
# pairing (integer work: must be bit-exact in the CUDA path)
# ----------------------------------------------------------------------------------------
PARTNER_PAD = -1  # label present in the other group but fewer cells there (precision 1, mu-term 0)
PARTNER_ABSENT = -2  # label absent from the other group's minibatch (mu 0, logvar 1)


def label_rank(labels: np.ndarray) -> np.ndarray:
    """rank of each row among the rows of its group with the same label, minibatch order
    (reference module/spVIPESmodule.py:685-701: label_count dict walk)."""
    labels = np.asarray(labels).reshape(-1)
    rank = np.zeros(labels.shape[0], dtype=np.int64)
    seen: Dict[int, int] = {}
    for i, l in enumerate(labels.tolist()):
        c = seen.get(l, 0)
        rank[i] = c
        seen[l] = c + 1
    return rank


def label_partners(labels_a: np.ndarray, labels_b: np.ndarray) -> np.ndarray:
    """For each row of group a: index of its partner row in group b, PARTNER_PAD or
    PARTNER_ABSENT (reference :599-659 + _poe2 padding :297-326)."""
    labels_a = np.asarray(labels_a).reshape(-1).astype(np.int64)
    labels_b = np.asarray(labels_b).reshape(-1).astype(np.int64)
    rank_a, rank_b = label_rank(labels_a), label_rank(labels_b)
    table = {(int(l), int(r)): j for j, (l, r) in enumerate(zip(labels_b, rank_b))}
    present = set(labels_b.tolist())
    out = np.empty(labels_a.shape[0], dtype=np.int64)
    for i, (l, r) in enumerate(zip(labels_a.tolist(), rank_a.tolist())):
        if l not in present:
            out[i] = PARTNER_ABSENT
        else:
            out[i] = table.get((l, r), PARTNER_PAD)
    return out


def sub_plan(plan: torch.Tensor, idx0, idx1) -> torch.Tensor:
    """T[idx0][:, idx1]  (reference _get_batch_transport_plans :474-482)."""
    i0 = torch.as_tensor(np.asarray(idx0).reshape(-1), dtype=torch.long)
    i1 = torch.as_tensor(np.asarray(idx1).reshape(-1), dtype=torch.long)
    return plan[i0][:, i1]


# ----------------------------------------------------------------------------------------
# the three product-of-experts variants.  Each returns per group (loc, logvar, scale_q)
# where scale_q is the scale of the posterior used for sampling AND for the KL term
# (already clamped in the OT modes, reference :274-276, 565-567; unclamped in label mode :712-714)
# ----------------------------------------------------------------------------------------
def _merge(mu_a, lv_a, mvp, ivp):
    """prior expert (precision 1) x own expert x partner expert, _poe2 :345-350."""
    var_a = torch.exp(lv_a)
    prec = 1.0 + 1.0 / var_a + ivp
    joint_var = 1.0 / prec
    mu = (mu_a / var_a + mvp) * joint_var
    return mu, torch.log(joint_var)


def poe_label(loc, lv, labels):
    out = []
    for a, b in ((0, 1), (1, 0)):
        part = label_partners(labels[a], labels[b])
        pt = torch.as_tensor(part)
        safe = pt.clamp(min=0)
        var_b = torch.exp(lv[b][safe])
        iv_b = 1.0 / var_b
        mv_b = loc[b][safe] / var_b
        is_real = (pt >= 0).unsqueeze(1)
        is_pad = (pt == PARTNER_PAD).unsqueeze(1)
        iv_abs = torch.full_like(iv_b, 1.0 / math.e) if iv_b.dtype == torch.float64 else 1.0 / torch.exp(torch.ones_like(iv_b))
        ivp = torch.where(is_real, iv_b, torch.where(is_pad, torch.ones_like(iv_b), iv_abs))
        mvp = torch.where(is_real, mv_b, torch.zeros_like(mv_b))
        mu, jlv = _merge(loc[a], lv[a], mvp, ivp)
        scale = torch.sqrt(torch.exp(jlv))  # _poe2 :358
        out.append((mu, jlv, scale, scale, part))
    return out


def poe_paired(loc, lv, sub):
    j_of_i = torch.argmax(sub, dim=1)
    i_of_j = torch.argmax(sub, dim=0)
    res = []
    for a, b, idx in ((0, 1, j_of_i), (1, 0, i_of_j)):
        var_b = torch.exp(lv[b][idx])
        mu, jlv = _merge(loc[a], lv[a], loc[b][idx] / var_b, 1.0 / var_b)
        scale = torch.exp(0.5 * jlv)  # :547-548
        res.append((mu, jlv, scale, scale.clamp(min=1e-6), idx.numpy().copy()))
    return res


def _normalize_plan(p):
    rs = p.sum(dim=1, keepdim=True).clamp(min=1e-10)
    return torch.where(p > 0, p / rs, p)


def poe_cluster(loc, lv, scale, sub, clabels):
    """reference _cluster_based_poe :184-280 (including the cross-indexing quirk :222,228,
    which requires B0 == B1)."""
    l0 = np.asarray(clabels[0]).reshape(-1).astype(np.int64)
    l1 = np.asarray(clabels[1]).reshape(-1).astype(np.int64)
    B0, B1 = l0.shape[0], l1.shape[0]
    outs = [[None] * B0, [None] * B1]  # not used; vector assembly below
    o_loc = [torch.zeros_like(loc[0]), torch.zeros_like(loc[1])]
    o_lv = [torch.zeros_like(lv[0]), torch.zeros_like(lv[1])]
    o_sc = [torch.zeros_like(scale[0]), torch.zeros_like(scale[1])]
    subT = sub.t()
    pieces = [[], []]  # (row indices, loc, lv, scale) per group
    for c in np.unique(np.concatenate([l0, l1])).tolist():
        m1 = torch.as_tensor(l0 == c)
        m2 = torch.as_tensor(l1 == c)
        r1 = torch.nonzero(m1).flatten()
        r2 = torch.nonzero(m2).flatten()
        n1, n2 = r1.numel(), r2.numel()
        if n1 > 0 and n2 > 0:
            p1 = _normalize_plan(sub[r1][:, r2])
            p2 = _normalize_plan(subT[r2][:, r1])
            a_loc, a_lv = p1 @ loc[0][r2], p1 @ lv[0][r2]  # group-0 stats indexed with group-1's mask (Q5)
            b_loc, b_lv = p2 @ loc[1][r1], p2 @ lv[1][r1]
            n = max(n1, n2)
            S = loc[0].shape[1]

            def pad(t, fill):
                if t.shape[0] == n:
                    return t
                return torch.cat([t, torch.full((n - t.shape[0], S), fill, dtype=t.dtype)], 0)

            va, vb = torch.exp(a_lv), torch.exp(b_lv)
            iv = 1.0 + pad(1.0 / va, 1.0) + pad(1.0 / vb, 1.0)
            mv = pad(a_loc / va, 0.0) + pad(b_loc / vb, 0.0)
            jv = 1.0 / iv
            jmu = mv * jv
            jlv = torch.log(jv)
            jsc = torch.sqrt(torch.exp(jlv))
            pieces[0].append((r1, jmu[:n1], jlv[:n1], jsc[:n1]))
            pieces[1].append((r2, jmu[:n2], jlv[:n2], jsc[:n2]))
        elif n1 > 0:
            pieces[0].append((r1, loc[0][r1], lv[0][r1], scale[0][r1]))
        elif n2 > 0:
            pieces[1].append((r2, loc[1][r2], lv[1][r2], scale[1][r2]))
    res = []
    for g in (0, 1):
        rows = torch.cat([p[0] for p in pieces[g]])
        inv = torch.empty_like(rows)
        inv[rows] = torch.arange(rows.numel())
        mu = torch.cat([p[1] for p in pieces[g]])[inv]
        jlv = torch.cat([p[2] for p in pieces[g]])[inv]
        sc = torch.cat([p[3] for p in pieces[g]])[inv]
        res.append((mu, jlv, sc, sc.clamp(min=1e-6), None))
    return res


# ----------------------------------------------------------------------------------------
# the step
# ----------------------------------------------------------------------------------------
def one_hot_batch(batch_index, n_batch: int, dtype):
    """reference nn/utils.py:9-13 (`one_hot`): [B] or [B, 1] batch codes -> [B, n_batch]; None when the covariate is not
    injected (n_batch <= 1: nn/networks.py:60-62 and scvi FCLayers zero the category count)."""
    if n_batch <= 1 or batch_index is None:
        return None
    idx = torch.as_tensor(np.asarray(batch_index) if not torch.is_tensor(batch_index) else batch_index).reshape(-1, 1).long()
    oh = torch.zeros(idx.shape[0], n_batch, dtype=dtype)
    oh.scatter_(1, idx, 1)
    return oh


def _with_cov(x, cov):
    return x if cov is None else torch.cat((x, cov), dim=-1)


def _relu(pre, key, gates, probe):
    """ReLU, with two test hooks: `probe` (dict) records the pre-activation under `key`; `gates` (dict) may hold a boolean
    mask for `key` that REPLACES the sign test (y = pre * gate).  A unit whose pre-activation is zero to within the rounding
    of the forward pass has an ill-defined gate: a checker comparing gradients fixes those units' gates to the decision the
    implementation under test took (tests/helpers.py: gate_consistent), every other unit keeps its own sign test."""
    if probe is not None:
        probe[key] = pre.detach()
    if gates is not None and key in gates:
        return pre * gates[key].to(pre.dtype)
    return F.relu(pre)


def encoder(sd, prefix, xl, training, drop_mask, new_stats, cov=None, gates=None, probe=None):
    """reference Encoder.forward nn/networks.py:110-125; `cov` = one-hot batch covariate appended to the fc1 input (:110-118)."""
    h = _relu(F.linear(_with_cov(xl, cov), sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]), prefix + ".fc1", gates, probe)
    h = _relu(F.linear(h, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"]), prefix + ".fc2", gates, probe)
    if training and drop_mask is not None:
        h = h * drop_mask  # mask already holds 0 or 1/(1-p)
    outs = []
    for head in ("mu_encoder", "lvar_encoder"):
        r = F.linear(h, sd[f"{prefix}.{head}.0.weight"], sd[f"{prefix}.{head}.0.bias"])
        k = f"{prefix}.{head}.1"
        outs.append(batch_norm(r, sd[k + ".weight"], sd[k + ".bias"], sd[k + ".running_mean"], sd[k + ".running_var"],
                               training, ENC_BN_EPS, ENC_BN_MOM, new_stats, k))
    loc, lv = outs
    return loc, lv, torch.exp(0.5 * lv)


def decoder(sd, g, z_private, z_shared, library, training, new_stats, cov=None, gates=None, probe=None):
    """reference LinearDecoderSPVIPE.forward nn/networks.py:314-325 + scvi FCLayers; `cov` = one-hot batch covariate that
    FCLayers appends to the input of each of the four one-layer nets (n_cat_list at nn/networks.py:203, 217, 245, 256)."""
    p = f"decoder_{g}"

    def fc(name, x, bn, eps=DEC_BN_EPS, mom=DEC_BN_MOM):
        k = f"{p}.{name}.fc_layers.Layer 0"
        y = F.linear(_with_cov(x, cov), sd[k + ".0.weight"], sd.get(k + ".0.bias"))
        if bn:
            y = batch_norm(y, sd[k + ".1.weight"], sd[k + ".1.bias"], sd[k + ".1.running_mean"], sd[k + ".1.running_var"],
                           training, eps, mom, new_stats, k + ".1")
        return y

    rate_p = torch.exp(library) * torch.softmax(fc("factor_regressor_private", z_private, True), dim=-1)
    rate_s = torch.exp(library) * torch.softmax(fc("factor_regressor_shared", z_shared, True), dim=-1)
    zz = torch.cat([z_private, z_shared], dim=1)
    hm = _relu(fc("sigmoid_decoder", zz, True), f"{p}.sigmoid_decoder", gates, probe)
    mix = fc("mixture", torch.cat([hm, zz], dim=-1), False)
    return rate_p, rate_s, mix


def step(sd: Dict[str, torch.Tensor], x: Sequence[torch.Tensor], *, mode: str, n_shared: int, n_private: int,
         eps_private: Sequence[torch.Tensor], eps_poe: Sequence[torch.Tensor],
         labels: Optional[Sequence] = None, sub: Optional[torch.Tensor] = None,
         drop_masks: Optional[Dict] = None, kl_weight: float = 1.0, training: bool = True,
         batch_index: Optional[Sequence] = None, n_batch: int = 0, gates: Optional[Dict] = None, probe: Optional[Dict] = None):
    """One forward pass of inference -> generative -> loss for both groups.

    x[g]: [B_g, G_g] counts of group g's OWN genes (any float dtype; the reference slices
    them out of the combined var axis at module/spVIPESmodule.py:428-430).
    mode: "label" | "paired" | "cluster".  labels: per-group int arrays (cell-type labels in
    label mode, processed_transport_labels in cluster mode).  sub: [B0, B1] sub-plan.
    batch_index / n_batch: per-group batch codes and the number of batches; with n_batch > 1 their one-hot is appended to
    the encoders' fc1 input and to the input of the four decoder nets (module/spVIPESmodule.py:133, 440-446, 748-756).
    gates / probe: test hooks of the ReLU gates, see _relu (None: plain ReLU everywhere, nothing recorded).
    Returns a dict of every quantity the parity gates name.
    """
    S, P = n_shared, n_private
    dt = x[0].dtype
    new_stats: Dict[str, torch.Tensor] = {}
    xl = [torch.log(1 + xs) for xs in x]  # :432-433
    lib = [torch.log(t.sum(1)).unsqueeze(1) for t in xl]  # :435 (sum of the log1p'd values)
    priv, shared = [], []
    cov = [one_hot_batch(batch_index[g] if batch_index is not None else None, n_batch, dt) for g in (0, 1)]
    for g in (0, 1):
        dm = drop_masks or {}
        priv.append(encoder(sd, f"encoder_{g}_private", xl[g], training, dm.get((g, "private")), new_stats, cov[g], gates, probe))
        shared.append(encoder(sd, f"encoder_{g}_shared", xl[g], training, dm.get((g, "shared")), new_stats, cov[g], gates, probe))
    s_loc = [shared[0][0], shared[1][0]]
    s_lv = [shared[0][1], shared[1][1]]
    s_sc = [shared[0][2], shared[1][2]]
    if mode == "label":
        poe = poe_label(s_loc, s_lv, labels)
    elif mode == "paired":
        poe = poe_paired(s_loc, s_lv, sub)
    elif mode == "cluster":
        poe = poe_cluster(s_loc, s_lv, s_sc, sub, labels)
    else:
        raise ValueError(mode)
    out = {"library": lib, "private_loc": [], "private_logvar": [], "private_log_z": [], "poe_loc": [],
           "poe_logvar": [], "poe_scale": [], "poe_log_z": [], "rec": [], "kl_private": [], "kl_poe": [],
           "partners": [], "shared_loc": s_loc, "shared_logvar": s_lv}
    total = 0.0
    for g in (0, 1):
        p_loc, p_lv, p_sc = priv[g]
        z_priv = p_loc + p_sc * eps_private[g]
        mu, jlv, sc, sc_q, part = poe[g]
        z_poe = mu + sc_q * eps_poe[g]
        c = torch.cat((z_priv, z_poe), dim=-1)  # :733  [private | poe]
        z_private_arg = c[:, S:S + P]  # :753  (quirk Q1)
        z_shared_arg = c[:, :S]  # :754
        rate_p, rate_s, mix = decoder(sd, g, z_private_arg, z_shared_arg, lib[g], training, new_stats, cov[g], gates, probe)
        theta = torch.exp(sd[f"px_r.{g}"])
        rec = -log_mixture_nb(xl[g], rate_p, rate_s, theta, mix).sum(-1)  # :820-824 (target = log1p counts, Q3)
        klp = kl_std_normal(p_loc, p_sc)
        klq = kl_std_normal(mu, sc_q)
        total = total + rec + kl_weight * klp + kl_weight * klq
        out["private_loc"].append(p_loc); out["private_logvar"].append(p_lv); out["private_log_z"].append(z_priv)
        out["poe_loc"].append(mu); out["poe_logvar"].append(jlv); out["poe_scale"].append(sc)
        out["poe_log_z"].append(z_poe); out["rec"].append(rec); out["kl_private"].append(klp)
        out["kl_poe"].append(klq); out["partners"].append(part)
    out["loss"] = torch.mean(total)  # :886-893
    out["new_stats"] = new_stats
    return out


# ----------------------------------------------------------------------------------------
# parameters / optimiser (scvi TrainingPlan defaults, restated: Adam lr 1e-3 eps 0.01 wd 1e-6)
# ----------------------------------------------------------------------------------------
def param_names(sd) -> List[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def default_state_dict(genes, n_hidden=128, n_shared=25, n_private=10, seed=0, decoder_hidden=256):
    """reference default initialisation as a plain state_dict (names of module/spVIPESmodule.py:118-120, 172-175): nn.Linear ->
    U(+-1/sqrt(fan_in)) for weight and bias, BatchNorm weight 1 / bias 0 / running stats 0 / 1, px_r ~ N(0, 1) (:115-117).
    For the CPU baseline of bench.py: the oracle then needs nothing from the product package."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def linear(name, n_out, n_in, bias=True):
        b = 1.0 / math.sqrt(n_in)
        sd[name + ".weight"] = (torch.rand(n_out, n_in, generator=gen) * 2 - 1) * b
        if bias:
            sd[name + ".bias"] = (torch.rand(n_out, generator=gen) * 2 - 1) * b

    def bn(name, n):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(n), torch.zeros(n)
        sd[name + ".running_mean"], sd[name + ".running_var"] = torch.zeros(n), torch.ones(n)

    kz = n_shared + n_private
    for g, G in enumerate(genes):
        for enc, n_out in ((f"encoder_{g}_private", n_private), (f"encoder_{g}_shared", n_shared)):
            linear(enc + ".fc1", n_hidden, G)
            linear(enc + ".fc2", n_hidden, n_hidden)
            for head in ("mu_encoder", "lvar_encoder"):
                linear(f"{enc}.{head}.0", n_out, n_hidden)
                bn(f"{enc}.{head}.1", n_out)
        fl = "fc_layers.Layer 0"
        for name, n_out, n_in, bias, has_bn in ((f"decoder_{g}.factor_regressor_private", G, n_private, False, True),
                                                (f"decoder_{g}.factor_regressor_shared", G, n_shared, False, True),
                                                (f"decoder_{g}.sigmoid_decoder", decoder_hidden, kz, True, True),
                                                (f"decoder_{g}.mixture", G, decoder_hidden + kz, True, False)):
            linear(f"{name}.{fl}.0", n_out, n_in, bias)
            if has_bn:
                bn(f"{name}.{fl}.1", n_out)
        sd[f"px_r.{g}"] = torch.randn(G, generator=gen)
    return sd


def adam_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=0.01, wd=1e-6):
    """torch.optim.Adam (L2 weight decay folded into the gradient), single tensor, step t>=1."""
    g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v
