"""As tools/nb_tile_trace.py, but inside the replayed CUDA graph of a whole training step: do both groups' launches of the
tiled NB forward kernel get all their CTAs resident at once?  Needs the -DNB_TRACE build (see nb_tile_trace.py)."""
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
mode, n_cells, genes, H, B, n_labels = bench.WORKLOADS["C2"]
lib = L.load()
clib = ctypes.CDLL(lib._name)
data = synth.make_counts((n_cells, n_cells), (genes, genes), n_labels, device=dev, seed=1234)
eng = StepEngine((genes, genes), H, bench.S_DIM, bench.P_DIM, 0.1, mode, device=dev, seed=0, precision="bf16")
init_params(eng, 0)
loop = TrainLoop(eng)
loop.set_epoch(1)
gen = torch.Generator(device=dev).manual_seed(5)
rows = [torch.randperm(n_cells, generator=gen, device=dev)[:B].to(torch.int32) for _ in (0, 1)]
batches = [GroupBatch(X=data.X[g], rows=rows[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
nG, nTB = (genes + 63) // 64, (B + 127) // 128
n = nTB * nG
trace = torch.zeros(2 * n * 8, dtype=torch.int64, device=dev)
clib.spv_debug_trace.argtypes = [ctypes.c_void_p]
clib.spv_debug_trace(trace.data_ptr())  # before the capture: the pointer is baked into the graph's kernel parameters
graph = loop.capture(batches)
for _ in range(20):
    graph.replay()
torch.cuda.synchronize()
clib.spv_debug_trace(None)
t = trace.view(2, n, 8).cpu()
t0 = t[:, :, 0].min().item()
for k in (0, 1):
    us = (t[k, :, :7] - t0).double() / 1e3
    sm = t[k, :, 7]
    ent = us[:, 0]
    srt = ent.sort().values
    print(f"launch slot {k}: first CTA enters at {ent.min():.2f} us, CTA #148 at {srt[147]:.2f}, #296 at {srt[295]:.2f}, last (#{n}) at {ent.max():.2f}; "
          f"kernel ends {us[:, 1].max():.2f}")
    per_sm = torch.bincount(sm.long(), minlength=148)
    early = torch.bincount(sm[ent < ent.min() + 3.0].long(), minlength=148)
    print(f"   CTAs per SM: min {per_sm.min()} max {per_sm.max()}; entered within 3 us of the first: {int((ent < ent.min() + 3.0).sum())} "
          f"(per SM max {early.max()})")
    late = (ent >= ent.min() + 3.0).nonzero().flatten().tolist()
    print("   late entries (us after the first):", ", ".join(f"{ent[i] - ent.min():.1f}" for i in late[:40]))
