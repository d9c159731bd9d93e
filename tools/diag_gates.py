"""diagnostic: are the large per-parameter gradient errors of the tensor-core mode vs the CPU oracle ReLU-gate flips?"""
import sys
import torch
sys.path.insert(0, ".")
import tests.test_gpu_shapes as T
from spvipes_b200.engine import GroupBatch, Noise, StepEngine
import torch.nn.functional as F

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
mode, B, G, H, NL = T.CONFIGS[cfg]
(data, rows, eps_p, eps_q, drop, plan), sd0, want, grads = T._case(cfg)
for prec in ("fp32", "bf16"):
    eng = StepEngine((G, G), H, T.S, T.P, 0.1, mode, "cuda", plan=plan, precision=prec)
    eng.load_state_dict(sd0); eng.set_kl_weight(0.25)
    noise = Noise([e.cuda() for e in eps_p], [e.cuda() for e in eps_q], [d.cuda() for d in drop])
    batches = []
    for g in (0, 1):
        r = rows[g].cuda()
        lab = data.labels[g] if mode == "label" else (data.labels[g][r.long()].contiguous() if mode == "cluster" else None)
        batches.append(GroupBatch(X=data.X[g], rows=r, labels=lab, labels_per_cell=mode == "label", idx=r))
    ws = eng.forward(batches, training=True, noise=noise); eng.backward(); torch.cuda.synchronize()
    print("==", cfg, prec)
    for g in (0, 1):
        x = torch.log1p(data.X[g].cpu().to(torch.int32)[rows[g].long()].float())
        for i, enc in enumerate(("private", "shared")):
            p = f"encoder_{g}_{enc}"
            pre1 = F.linear(x, sd0[p + ".fc1.weight"], sd0[p + ".fc1.bias"])
            h1 = F.relu(pre1)
            pre2 = F.linear(h1, sd0[p + ".fc2.weight"], sd0[p + ".fc2.bias"])
            e1 = ws[g].h1[:, i * H:(i + 1) * H].cpu(); e2 = ws[g].h2[:, i * H:(i + 1) * H].cpu()
            m = drop[g][:, i * H:(i + 1) * H] > 0
            f1 = ((e1 > 0) != (pre1 > 0)); f2 = (((e2 > 0) != (pre2 > 0)) & m)
            print(f"  {p}: gate flips fc1 {int(f1.sum())} (|pre| there {pre1[f1].abs().tolist()[:4]}), fc2 {int(f2.sum())} (|pre| {pre2[f2].abs().tolist()[:4]}); "
                  f"units with |pre| < 1e-5: {int((pre1.abs() < 1e-5).sum())} / {int((pre2.abs() < 1e-5).sum())}")
    got = {k: v.cpu() for k, v in eng.grad_dict().items()}
    errs = sorted(((float((got[k] - w).abs().max() / (w.abs().max() + 1e-30)), k) for k, w in grads.items()
                   if not k.endswith(("mu_encoder.0.bias", "lvar_encoder.0.bias", "sigmoid_decoder.fc_layers.Layer 0.0.bias"))), reverse=True)
    for e, k in errs[:6]:
        d = (got[k] - grads[k]).abs()
        rowmax = d.reshape(d.shape[0], -1).max(1).values if d.dim() > 1 else d
        print(f"  {k:60s} {e:.2e}   rows with error > 20% of the worst: {int((rowmax > 0.2 * rowmax.max()).sum())} of {rowmax.numel()}")
