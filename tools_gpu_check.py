"""diagnostic (not a test): print per-output errors of the CUDA path against every golden fixture.
usage: python tools_gpu_check.py [--precision bf16] [fixture names...]"""
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.helpers import Golden, golden_names, relerr  # noqa: E402
from tests.gpu_helpers import engine_from_golden, engine_outputs  # noqa: E402

KEYS = ("library", "private_loc", "private_logvar", "shared_loc", "shared_logvar", "private_log_z", "poe_loc", "poe_logvar",
        "poe_scale", "poe_log_z", "kl_private", "kl_poe", "rec")
argv = sys.argv[1:]
precision = "fp32"
if "--precision" in argv:
    i = argv.index("--precision")
    precision = argv[i + 1]
    del argv[i:i + 2]
names = argv or golden_names()
for name in names:
    try:
        gd = Golden(name)
        eng, batches, noise = engine_from_golden(gd, precision=precision)
        ws = eng.forward(batches, training=gd.training, noise=noise)
        torch.cuda.synchronize()
        out = engine_outputs(eng, ws)
        print(f"== {name} mode={gd.mode} {precision} loss got={float(out['loss']):.6f} want={float(gd.out['loss']):.6f} "
              f"rel={relerr(out['loss'], gd.out['loss']):.2e}")
        for k in KEYS:
            print(f"   {k:16s} " + " ".join(f"{relerr(out[k][g].reshape(-1), gd.out[f'{k}{g}'].reshape(-1)):.2e}" for g in (0, 1)))
        if gd.mode in ("label", "paired"):
            print("   partners equal:", [bool(np.array_equal(out["partners"][g], gd.out[f"partner{g}"])) for g in (0, 1)])
        if gd.training:
            eng.backward()
            torch.cuda.synchronize()
            scale = max(float(v.abs().max()) for v in gd.grads.values())
            rows = []
            for k, wv in gd.grads.items():
                gv = eng.grad_dict()[k].cpu()
                dd = float((gv.double().reshape(-1) - wv.double().reshape(-1)).abs().max())
                rows.append((dd / (float(wv.abs().max()) + 1e-30), dd / scale, float(wv.abs().max()), k))
            gg = torch.cat([eng.grad_dict()[k].cpu().double().reshape(-1) for k in gd.grads])
            ww = torch.cat([v.double().reshape(-1) for v in gd.grads.values()])
            print(f"   overall grad cosine {float(torch.dot(gg, ww) / (gg.norm() * ww.norm())):.6f}  rel l2 err {float((gg - ww).norm() / ww.norm()):.3e}")
            cos = []
            for k, wv in gd.grads.items():
                gv = eng.grad_dict()[k].cpu().double().reshape(-1)
                w1 = wv.double().reshape(-1)
                if float(w1.abs().max()) > 1e-4:
                    cos.append((float(torch.dot(gv, w1) / (gv.norm() * w1.norm() + 1e-30)), k))
            cos.sort()
            print("   lowest per-parameter cosines:", [(round(c, 4), k) for c, k in cos[:5]])
            rows.sort(reverse=True)
            for r in rows[:48]:
                print(f"   grad rel={r[0]:.2e} relglobal={r[1]:.2e} max|g|={r[2]:.2e} {r[3]}")
    except Exception:
        traceback.print_exc()
