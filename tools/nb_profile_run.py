"""small driver for ncu: a few eager training steps at a BASELINE minibatch shape (C5 by default) on a reduced resident matrix.
    python tools/nb_profile_run.py [C5|C2] && ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:"nb_tc_(fwd|bwd|train)" -c 4 -o gpurun_out/prof python tools/nb_profile_run.py"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from spvipes_b200 import synth
from spvipes_b200.engine import GroupBatch, StepEngine
from spvipes_b200.trainer import TrainLoop, init_params

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
mode, n_cells, G, H, B, NL, _ = bench.WORKLOADS[wl]
N = min(n_cells, 16 * B)
data = synth.make_counts((N, N), (G, G), NL, device="cuda", seed=1234)
eng = StepEngine((G, G), H, 25, 10, 0.1, mode, "cuda", seed=0, precision="bf16")
init_params(eng, 0)
eng.parallel_groups = False
loop = TrainLoop(eng)
loop.set_epoch(1)
gen = torch.Generator(device="cuda").manual_seed(5)
for s in range(3):
    if s == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: the third step only
    rows = [torch.randperm(N, generator=gen, device="cuda")[:B].to(torch.int32) for _ in (0, 1)]
    loop.step([GroupBatch(X=data.X[g], rows=rows[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)])
torch.cuda.synchronize()
print("ok", float(eng.loss_out[0]))
