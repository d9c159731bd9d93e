"""Per-CTA timestamps of the tiled NB forward kernel (diagnostic).  Needs a library built with the stamps:
    SPV_NVCC_EXTRA="-DNB_TRACE" python -m spvipes_b200.build --force
    python tools/nb_tile_trace.py
Stamps per CTA (thread 64 = first epilogue thread): 0 kernel entry, 2 count gather issued, 3 accumulators complete, 4 counts
staged in shared memory, 5 this warp's epilogue done, 1 every warp done (after the closing barrier); thread 32 (MMA warp):
6 tensor memory granted; slot 7 = SM id."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spvipes_b200 import _lib as L, synth  # noqa: E402
from spvipes_b200.engine import GroupBatch, StepEngine  # noqa: E402
from spvipes_b200.trainer import TrainLoop, init_params  # noqa: E402

dev = torch.device("cuda", 0)
WL = os.environ.get("WL", "C2")
mode, n_cells, genes, H, B, n_labels = bench.WORKLOADS[WL][:6]
n_cells = min(n_cells, 60000)
lib = L.load()
clib = ctypes.CDLL(lib._name)
data = synth.make_counts((n_cells, n_cells), (genes, genes), n_labels, device=dev, seed=1234)
eng = StepEngine((genes, genes), H, bench.S_DIM, bench.P_DIM, 0.1, mode, device=dev, seed=0, precision="bf16")
eng.parallel_groups = False  # one group's kernel at a time, so that the stamps of a launch are not mixed with the other's
init_params(eng, 0)
loop = TrainLoop(eng)
loop.set_epoch(1)
gen = torch.Generator(device=dev).manual_seed(5)
rows = [torch.randperm(n_cells, generator=gen, device=dev)[:B].to(torch.int32) for _ in (0, 1)]
batches = [GroupBatch(X=data.X[g], rows=rows[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
for _ in range(3):
    loop.step(batches)
torch.cuda.synchronize()
nG, nTB = (genes + 63) // 64, (B + 127) // 128
trace = torch.zeros(nTB * nG * 8, dtype=torch.int64, device=dev)
clib.spv_debug_trace.argtypes = [ctypes.c_void_p]
clib.spv_debug_trace(trace.data_ptr())
eng.forward(batches, training=True)  # the second group's launch overwrites the first's stamps
torch.cuda.synchronize()
clib.spv_debug_trace(None)
t = trace.view(nTB * nG, 8).cpu()
t0 = t[:, 0].min().item()
us = (t[:, :7] - t0).double() / 1e3
names = ["entry", "setup done", "gather issued", "acc complete", "counts staged", "epilogue done"]
print(f"{nTB * nG} CTAs; kernel span {us[:, 5].max():.2f} us")
first = us[:, 0] < 1.0
print(f"first wave: {int(first.sum())} CTAs; later: {int((~first).sum())}")
for sel, tag in ((first, "first wave"), (~first, "later CTAs")):
    if sel.sum() == 0:
        continue
    u = us[sel]
    print(tag)
    print("   entry at        : median %6.2f  min %6.2f  max %6.2f" % (u[:, 0].median(), u[:, 0].min(), u[:, 0].max()))
    for i in range(3, 6):
        d = u[:, i] - u[:, i - 1]
        print("   %-15s : median %6.2f  min %6.2f  max %6.2f us after the previous stamp" % (names[i], d.median(), d.min(), d.max()))
    d = u[:, 5] - u[:, 0]
    print("   CTA lifetime    : median %6.2f  min %6.2f  max %6.2f" % (d.median(), d.min(), d.max()))
sm = t[:, 7]
per_sm = torch.bincount(sm.clamp(min=0).long(), minlength=148)
print("CTAs per SM: min %d max %d; SMs with 3+: %d" % (per_sm.min(), per_sm.max(), int((per_sm >= 3).sum())))
shown = 0
for s in range(148):
    ids = (sm == s).nonzero().flatten().tolist()
    if len(ids) >= 3 and shown < 3:
        shown += 1
        print(f"SM {s}:")
        for i in sorted(ids, key=lambda i: us[i, 0].item()):
            print("   cta %4d: entry %6.2f  gather %6.2f  acc %6.2f  staged %6.2f  done %6.2f   (epilogue %5.2f us)"
                  % (i, us[i, 0], us[i, 2], us[i, 3], us[i, 4], us[i, 5], us[i, 5] - us[i, 4]))
for s in range(148):
    ids = (sm == s).nonzero().flatten().tolist()
    if len(ids) == 2:
        print(f"SM {s} (two CTAs):")
        for i in ids:
            print("   cta %4d: entry %6.2f  gather %6.2f  acc %6.2f  staged %6.2f  done %6.2f   (epilogue %5.2f us)"
                  % (i, us[i, 0], us[i, 2], us[i, 3], us[i, 4], us[i, 5], us[i, 5] - us[i, 4]))
        break

if WL != "C2":  # steady state: every CTA of one SM in entry order
    s0 = int(sm[0])
    ids = sorted((sm == s0).nonzero().flatten().tolist(), key=lambda i: us[i, 0].item())
    print(f"SM {s0}: {len(ids)} CTAs")
    for i in ids[:24]:
        print("   cta %5d: entry %7.2f  gather %7.2f  tmem %7.2f  acc %7.2f  warp done %7.2f  cta done %7.2f   (epilogue %5.2f us)"
              % (i, us[i, 0], us[i, 2], us[i, 6], us[i, 3], us[i, 5], us[i, 1], us[i, 5] - us[i, 4]))
