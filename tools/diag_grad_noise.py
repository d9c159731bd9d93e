"""diagnostic: where does the tensor-core mode's gradient error come from?  fp32 engine vs bf16 engine on identical inputs
(same weights, rows, explicit noise), relative max-norm difference of the backward intermediates, per config shape."""
import sys
import torch
sys.path.insert(0, ".")
from spvipes_b200 import synth
from spvipes_b200.engine import GroupBatch, Noise, StepEngine
from spvipes_b200.trainer import init_params

S, P = 25, 10


def run(mode, B, G, H, NL=10):
    data = synth.make_counts((B, B), (G, G), NL, device="cuda", seed=5)
    plan = synth.make_plan(B, B, data.labels[0], data.labels[1], NL, device="cuda") if mode != "label" else None
    gen = torch.Generator().manual_seed(1)
    noise = Noise([torch.randn(B, P, generator=gen).cuda() for _ in (0, 1)], [torch.randn(B, S, generator=gen).cuda() for _ in (0, 1)],
                  [((torch.rand(B, 2 * H, generator=gen) < 0.9).float() / 0.9).cuda() for _ in (0, 1)])
    idx = [torch.arange(B, dtype=torch.int32, device="cuda") for _ in (0, 1)]
    res = {}
    for prec in ("fp32", "bf16"):
        eng = StepEngine((G, G), H, S, P, 0.1, mode, "cuda", plan=plan, precision=prec)
        init_params(eng, 3)
        eng.set_kl_weight(0.25)
        bt = [GroupBatch(X=data.X[g], labels=data.labels[g] if mode != "paired" else None, idx=idx[g]) for g in (0, 1)]
        ws = eng.forward(bt, training=True, noise=noise)
        eng.backward()
        torch.cuda.synchronize()
        w = ws[0]
        res[prec] = {k: getattr(w, k).clone() for k in ("h1", "stats", "zpoe", "rec", "damix", "dzraw", "dzz", "dstats", "dr", "dh2", "dh1", "dah")}
        res[prec]["grads"] = {k: v.clone() for k, v in eng.grad_dict().items()}
    print(f"== {mode} B={B} G={G} H={H}")
    for k in res["fp32"]:
        if k == "grads":
            continue
        a, b = res["bf16"][k].double(), res["fp32"][k].double()
        print(f"  {k:8s} max-norm rel {float((a - b).abs().max() / b.abs().max()):.2e}   l2 rel {float((a - b).norm() / b.norm()):.2e}"
              f"   |mean|/rms {float(b.mean(0).abs().max() / b.pow(2).mean().sqrt()):.2e}")
    errs = []
    for k, b in res["fp32"]["grads"].items():
        if k.endswith(("mu_encoder.0.bias", "lvar_encoder.0.bias", "sigmoid_decoder.fc_layers.Layer 0.0.bias")):
            continue  # analytically zero (a bias in front of a training-mode BatchNorm)
        a = res["bf16"]["grads"][k]
        errs.append((float((a - b).abs().max() / (b.abs().max() + 1e-30)), float((a - b).norm() / (b.norm() + 1e-30)), k, float(b.abs().max())))
    for e in sorted(errs, reverse=True)[:16]:
        print(f"  grad {e[2]:55s} max-norm rel {e[0]:.2e}  l2 rel {e[1]:.2e}  max|g| {e[3]:.2e}")


run("label", 512, 5000, 128)
run("paired", 1024, 10000, 128)
