import sys, os
import numpy as np, pandas as pd, torch
sys.path.insert(0, ".")
from spvipes_b200.engine import Noise
from spvipes_b200.model import GroupedData, prepare_adatas, spVIPES
name = sys.argv[1] if len(sys.argv) > 1 else "latent_label_ragged"
z = np.load(os.path.join("tests", "golden_latent", name + ".npz"))
mode = str(z["meta_mode"]); n0, n1, G0, G1, H, S, P, nl, bs = (int(v) for v in z["meta_dims"])
ads = {}
for gi, key in enumerate(("a_first", "b_second")):
    obs = pd.DataFrame({"cell_type": [f"t{int(v)}" for v in z[f"labels{gi}"]]})
    if mode == "cluster": obs["processed_transport_labels"] = z[f"labels{gi}"]
    ads[key] = GroupedData(X=z[f"x{gi}"].astype(np.float32), obs=obs, var_names=[f"g{j}" for j in range((G0, G1)[gi])])
adata = prepare_adatas(ads)
if mode != "label": adata.uns["transport_plan"] = z["plan"]
spVIPES.setup_anndata(adata, groups_key="groups", label_key="cell_type" if mode == "label" else None,
                      transport_plan_key="transport_plan" if mode != "label" else None, match_clusters=mode == "cluster")
model = spVIPES(adata, n_hidden=H, n_dimensions_shared=S, n_dimensions_private=P, dropout_rate=0.1, precision="fp32")
sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
print(model.module.load_state_dict(sd, strict=True))
def noise(k, B0, B1):
    g = torch.Generator().manual_seed(1000 + k)
    ep = [torch.randn(B, P, generator=g) for B in (B0, B1)]; eq = [torch.randn(B, S, generator=g) for B in (B0, B1)]
    return Noise([e.cuda() for e in ep], [e.cuda() for e in eq], None)
gil = [list(ix) for ix in adata.uns["groups_obs_indices"]]
lat = model.get_latent_representation(gil, batch_size=bs, _noise_fn=noise)
for key, got in (("private0", lat["private"][0]), ("private1", lat["private"][1]), ("shared0", lat["shared"][0]), ("shared1", lat["shared"][1])):
    d = np.abs(got - z[key]).max(1) / np.abs(z[key]).max()
    print(key, "max", d.max(), "rows > 1e-3:", np.nonzero(d > 1e-3)[0][:40].tolist())
# engine vs oracle on the first minibatch, eval mode
from oracle import restatement as rs
from spvipes_b200.engine import GroupBatch
eng = model.module.engine
B = bs
x = [torch.from_numpy(z[f"x{g}"][:B].astype(np.float32)) for g in (0, 1)]
labels = [z[f"labels{g}"][:B] for g in (0, 1)]
g0 = torch.Generator().manual_seed(1000)
ep = [torch.randn(B, P, generator=g0) for _ in (0, 1)]; eq = [torch.randn(B, S, generator=g0) for _ in (0, 1)]
sub = torch.from_numpy(z["plan"])[:B, :B] if mode != "label" else None
o = rs.step({k: v.clone() for k, v in sd.items()}, x, mode=mode, n_shared=S, n_private=P, eps_private=ep, eps_poe=eq, labels=labels if mode != "paired" else None, sub=sub, training=False)
print("oracle private_log_z[0][0,:4]", o["private_log_z"][0][0, :4].tolist(), " fixture", z["private0"][0, :4].tolist(), " engine", lat["private"][0][0, :4].tolist())
print("oracle private_loc[0][0,:4]", o["private_loc"][0][0, :4].tolist())
d = model._device_data
bt = [GroupBatch(X=d[g]["X"], rows=torch.arange(B, dtype=torch.int32, device="cuda"), labels=d[g]["labels"][:B].contiguous() if d[g]["labels"] is not None else None, idx=d[g]["idx"][:B].contiguous(), B=B) for g in (0, 1)]
ws = eng.forward(bt, training=False, noise=noise(0, B, B), with_grad=False, decode=False)
torch.cuda.synchronize()
print("engine stats loc[0,:4]", ws[0].stats[0, :4].tolist(), "zpriv", ws[0].zpriv[0, :4].tolist())
print("X row0 engine", d[0]["X"][0, :8].tolist(), "fixture", z["x0"][0, :8].tolist())
for k in ("encoder_0_private.mu_encoder.1.running_mean", "encoder_0_private.mu_encoder.1.running_var", "encoder_0_private.fc1.weight", "encoder_0_private.mu_encoder.1.weight"):
    a = eng.state_dict()[k].cpu(); b = sd[k]
    print(k, "engine==fixture:", float((a - b).abs().max()))
from spvipes_b200.engine import StepEngine
e2 = StepEngine((G0, G1), H, S, P, 0.1, mode, "cuda", plan=torch.from_numpy(z["plan"]).cuda() if mode != "label" else None)
e2.load_state_dict(sd)
ws2 = e2.forward(bt, training=False, noise=noise(0, B, B), with_grad=False, decode=False)
torch.cuda.synchronize()
print("fresh engine loc", ws2[0].stats[0, :4].tolist())
print("h1 diff module-engine vs fresh", float((ws[0].h1 - ws2[0].h1).abs().max()), "r diff", float((ws[0].r - ws2[0].r).abs().max()))
xl = torch.log1p(x[0]); h = torch.relu(xl @ sd["encoder_0_private.fc1.weight"].t() + sd["encoder_0_private.fc1.bias"])
print("oracle h1[0,:4]", h[0, :4].tolist(), "engine", ws2[0].h1[0, :4].tolist())
h2 = torch.relu(h @ sd["encoder_0_private.fc2.weight"].t() + sd["encoder_0_private.fc2.bias"])
print("oracle h2[0,:4]", h2[0, :4].tolist(), "engine", ws2[0].h2[0, :4].tolist())
r = h2 @ sd["encoder_0_private.mu_encoder.0.weight"].t() + sd["encoder_0_private.mu_encoder.0.bias"]
print("oracle r[0,:4]", r[0, :4].tolist(), "engine", ws2[0].r[0, :4].tolist())
