"""Per-unit timestamps of the persistent NB forward kernel (diagnostic).  Needs a library built with the trace code and the
opt-in kernel selected:
    SPV_NVCC_EXTRA="-DPTC_TRACE" python -m spvipes_b200.build --force
    SPV_NVCC_EXTRA="-DPTC_TRACE" SPV_NB_PERSISTENT=1 python tools/ptc_trace.py
(-DPTC_EXP=2 replaces every MUFU by an FFMA: one of the timing experiments of profiles/r1_nb_persistent_notes.md; the
count-gather and no-math variants were one-off edits and are not in the tree)."""
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
eng.parallel_groups = len(sys.argv) > 1 and sys.argv[1] == "par"
init_params(eng, 0)
loop = TrainLoop(eng)
loop.set_epoch(1)
gen = torch.Generator(device=dev).manual_seed(5)
rows = [torch.randperm(n_cells, generator=gen, device=dev)[:B].to(torch.int32) for _ in (0, 1)]
batches = [GroupBatch(X=data.X[g], rows=rows[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
for _ in range(3):
    loop.step(batches)
torch.cuda.synchronize()
trace = torch.zeros(148 * 2 * 16 * 4, dtype=torch.int64, device=dev)
clib.spv_debug_trace.argtypes = [ctypes.c_void_p]
clib.spv_debug_trace(trace.data_ptr())
eng.forward(batches, training=True)
torch.cuda.synchronize()
clib.spv_debug_trace(None)
t = trace.view(148, 2, 16, 4).cpu()
t0 = t[t > 0].min().item()
for cta in (0, 1, 73, 147):
    for eg in (0, 1):
        print(f"cta {cta} group {eg}: (start, after const barrier, acc ready | next...) in us relative to the earliest stamp; last = group done")
        for n in range(16):
            s = t[cta, eg, n]
            if s[0] > 0:
                print("   unit %2d: start %7.2f  barrier +%5.2f  acc wait +%5.2f" % (n, (s[0] - t0) / 1e3, (s[1] - s[0]) / 1e3, (s[2] - s[1]) / 1e3))
        print("   done %7.2f" % ((t[cta, eg, 15, 3] - t0) / 1e3))
