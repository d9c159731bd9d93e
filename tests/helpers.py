"""Shared test helpers: golden-fixture loading, oracle invocation, comparison metrics."""
import glob
import os

import numpy as np
import torch

from oracle import restatement as rs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# Linear biases that feed straight into a training-mode BatchNorm: their gradient is
# analytically zero, both sides produce rounding noise only.
ZERO_GRAD_SUFFIXES = ("mu_encoder.0.bias", "lvar_encoder.0.bias", "sigmoid_decoder.fc_layers.Layer 0.0.bias")


def golden_names(training=None):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if training is None:
        return names
    return [n for n in names if n.endswith("_eval") != training]


# fixtures of functionality the CUDA path does not cover yet (batch covariates): oracle-only, not enumerated by the GPU tests
GOLDEN_NEXT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_next")


def golden_next_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_NEXT_DIR, "*.npz")))


class Golden:
    def __init__(self, name, directory=GOLDEN_DIR):
        z = np.load(os.path.join(directory, name + ".npz"))
        self.name = name
        self.mode = str(z["meta_mode"])
        self.B, self.G0, self.G1, self.H, self.S, self.P, self.n_labels, self.N = (int(v) for v in z["meta_dims"])
        self.dropout = float(z["meta_dropout"])
        self.kl_weight = float(z["meta_kl_weight"])
        self.training = bool(z["meta_training"])
        self.n_batch = int(z["meta_n_batch"]) if "meta_n_batch" in z.files else 0
        self.batch = [z[f"batch{g}"] for g in (0, 1)] if self.n_batch > 1 else None
        self.plan = torch.from_numpy(z["plan"])
        self.x = [torch.from_numpy(z[f"x{g}"].astype(np.int32)) for g in (0, 1)]
        self.idx = [z[f"idx{g}"] for g in (0, 1)]
        self.labels = [z[f"labels{g}"] for g in (0, 1)]
        self.eps_private = [torch.from_numpy(z[f"eps_private{g}"]) for g in (0, 1)]
        self.eps_poe = [torch.from_numpy(z[f"eps_poe{g}"]) for g in (0, 1)]
        self.keep = {(g, k): torch.from_numpy(z[f"keep_{g}_{k}"]) for g in (0, 1) for k in ("private", "shared")}
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        self.grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
        self.after = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("after/")}
        self.out = {}
        for k in z.files:
            if k.startswith("out_"):
                self.out[k[4:]] = z[k]

    def drop_masks(self, dtype=torch.float32):
        if self.dropout <= 0:
            return None
        keep = 1.0 - self.dropout
        return {k: v.to(dtype) / keep for k, v in self.keep.items()}

    def sub(self, dtype=torch.float32):
        if self.mode == "label":
            return None
        return rs.sub_plan(self.plan, self.idx[0], self.idx[1]).to(dtype)


def run_oracle(gd: Golden, dtype=torch.float32, backward=True, gates=None, probe=None):
    sd = {}
    for k, v in gd.sd.items():
        if v.is_floating_point():
            v = v.to(dtype).clone()
            if "running" not in k:
                v.requires_grad_(backward and gd.training)
        sd[k] = v
    out = rs.step(sd, [t.to(dtype) for t in gd.x], mode=gd.mode, n_shared=gd.S, n_private=gd.P,
                  eps_private=[e.to(dtype) for e in gd.eps_private], eps_poe=[e.to(dtype) for e in gd.eps_poe],
                  labels=gd.labels, sub=gd.sub(dtype), drop_masks=gd.drop_masks(dtype), kl_weight=gd.kl_weight,
                  training=gd.training, batch_index=gd.batch, n_batch=gd.n_batch, gates=gates, probe=probe)
    grads = None
    if backward and gd.training:
        out["loss"].backward()
        grads = {k: sd[k].grad for k in rs.param_names(sd)}
    return out, grads, sd


def relerr(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def grad_errors(got, want):
    """max relative error over parameters (each normalised by its own max |grad|); the
    analytically-zero pre-BatchNorm biases are normalised by the global gradient scale."""
    scale = max(float(v.abs().max()) for v in want.values())
    worst, where = 0.0, None
    for k, w in want.items():
        g = got[k]
        assert g is not None, f"missing gradient for {k}"
        d = float((g.double().reshape(-1) - w.double().reshape(-1)).abs().max())
        denom = scale if k.endswith(ZERO_GRAD_SUFFIXES) else float(w.abs().max()) + 1e-30
        if d / denom > worst:
            worst, where = d / denom, k
    return worst, where


GATE_AMBIGUITY = 1e-4  # relative to the layer's largest pre-activation: below the 1e-3 forward gate on the latents


def engine_gates(eng, ws, n_hidden):
    """the ReLU gate decisions the CUDA path took, keyed like oracle.restatement._relu: encoders' fc1 / fc2 and the decoder's
    hidden layer.  h2 is stored after dropout, so only units the mask kept are informative (the others carry no gradient)."""
    H = n_hidden
    out = {}
    for g, w in enumerate(ws):
        for i, enc in enumerate(("private", "shared")):
            out[f"encoder_{g}_{enc}.fc1"] = (w.h1[:, i * H:(i + 1) * H] > 0).cpu()
            out[f"encoder_{g}_{enc}.fc2"] = (w.h2[:, i * H:(i + 1) * H] != 0).cpu()
        out[f"decoder_{g}.sigmoid_decoder"] = (w.amix[:, :256] > 0).cpu()
    return out


def gate_consistent(probe, eng_gates, drop_masks=None):
    """gates for a second oracle pass: the oracle's own sign test everywhere, except on units whose pre-activation is zero to
    within GATE_AMBIGUITY (relative), where the implementation's decision is taken.  Asserts that the two sides disagree ONLY
    on such units, and returns (gates, number of units that were switched)."""
    gates, switched = {}, 0
    for key, pre in probe.items():
        own = pre > 0
        theirs = eng_gates[key]
        if key.endswith(".fc2") and drop_masks is not None:  # dropped units: h2 == 0 whatever the gate was
            g, enc = int(key.split("_")[1]), key.split("_")[2].split(".")[0]
            kept = drop_masks[(g, enc)] > 0
            theirs = torch.where(kept, theirs, own)
        diff = own != theirs
        if bool(diff.any()):
            worst = float(pre[diff].abs().max() / pre.abs().max())
            assert worst < GATE_AMBIGUITY, (key, int(diff.sum()), worst)
            switched += int(diff.sum())
        gates[key] = torch.where(diff, theirs, own)
    return gates, switched
