"""micro-benchmark of the fused first-layer kernels: compact minibatch matrix vs rows scattered over a large resident matrix"""
import sys
import torch
sys.path.insert(0, ".")
from spvipes_b200 import _lib as L
lib = L.load()
B, G, N = 2048, 20000, 256
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
W = (torch.rand(N, G, generator=g, device="cuda") * 2 - 1) * 0.05
Wh = torch.zeros(N, G, device="cuda", dtype=torch.bfloat16); Wl = torch.zeros_like(Wh)
L.check(lib.spv_to_bf16_split(W.data_ptr(), G, Wh.data_ptr(), Wl.data_ptr(), G, N, G, st), "split")
d = torch.randn(B, N, generator=g, device="cuda")
dh = torch.zeros(B, N, device="cuda", dtype=torch.bfloat16); dl = torch.zeros_like(dh)
L.check(lib.spv_to_bf16_split(d.data_ptr(), N, dh.data_ptr(), dl.data_ptr(), N, B, N, st), "split")
h1 = torch.empty(B, N, device="cuda"); ws = torch.empty(8 * B * N, device="cuda"); dW = torch.empty(N, G, device="cuda")
for name, nrows in (("compact [B, G]", B), ("scattered over 100k rows", 100000), ("scattered over 400k rows", 400000)):
    X = torch.randint(0, 4, (nrows, G), generator=g, device="cuda", dtype=torch.int32).to(torch.uint16)
    rows = (torch.randperm(nrows, generator=g, device="cuda")[:B].to(torch.int32) if nrows > B else None)
    rp = rows.data_ptr() if rows is not None else None
    for fn, label in ((lambda: lib.spv_enc_fc1_fwd(X.data_ptr(), G, rp, Wh.data_ptr(), Wl.data_ptr(), G, h1.data_ptr(), N, B, N, G, None, 1, 0, 4, ws.data_ptr(), st), "fwd"),
                      (lambda: lib.spv_enc_fc1_dw(X.data_ptr(), G, rp, dh.data_ptr(), dl.data_ptr(), N, dW.data_ptr(), G, B, N, G, st), "dw ")):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:28s} {label} {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us")
    del X
