"""host-side cost of the plugin call per step (C2 shape): where does the time go between the two graph replays?"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from spvipes_b200 import synth
from spvipes_b200.module import spVIPESmodule

B, G, H, NL = 512, 5000, 128, 10
data = synth.make_counts((8 * B, 8 * B), (G, G), NL, device="cuda", seed=1)
torch.manual_seed(0)
m = spVIPESmodule(groups_lengths={0: G, 1: G}, groups_obs_names=[None, None], groups_var_names={0: None, 1: None},
                  groups_obs_indices=[None, None], groups_var_indices=[np.arange(G), np.arange(G, 2 * G)], use_labels=True, n_labels=NL,
                  precision="bf16")
m.train()
kind = sys.argv[1] if len(sys.argv) > 1 else "foreach"
if kind == "flat":
    from spvipes_b200.optim import FlatAdam
    opt = FlatAdam(m)
else:
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, eps=0.01, weight_decay=1e-6, fused=kind == "fused")
host = []
for s in range(4):
    batch = []
    for g in (0, 1):
        r = torch.randperm(8 * B)[:B]
        X = torch.zeros(B, 2 * G).pin_memory()
        X[:, g * G:(g + 1) * G] = data.X[g].cpu().to(torch.int32)[r].float()
        batch.append({"X": X, "batch": torch.zeros(B, 1), "groups": torch.full((B, 1), float(g)), "indices": r.float().reshape(-1, 1).pin_memory(),
                      "labels": data.labels[g].cpu()[r].float().reshape(-1, 1).pin_memory()})
    host.append(tuple(batch))
T = {"zero": 0, "fwd": 0, "bwd": 0, "opt": 0}
N = 200
for s in range(N + 10):
    if s == 10:
        torch.cuda.synchronize(); T = {k: 0 for k in T}; t_all = time.perf_counter()
    t0 = time.perf_counter(); opt.zero_grad(set_to_none=True)
    t1 = time.perf_counter(); _, _, lo = m(host[s % 4], loss_kwargs={"kl_weight": 0.5})
    t2 = time.perf_counter(); lo.loss.backward()
    t3 = time.perf_counter(); opt.step()
    t4 = time.perf_counter()
    T["zero"] += t1 - t0; T["fwd"] += t2 - t1; T["bwd"] += t3 - t2; T["opt"] += t4 - t3
torch.cuda.synchronize()
tot = time.perf_counter() - t_all
print(kind, {k: round(v / N * 1e3, 3) for k, v in T.items()}, "ms host per step; wall per step", round(tot / N * 1e3, 3))
