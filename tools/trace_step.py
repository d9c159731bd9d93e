"""Kernel timeline of ONE graph-replayed training step (CUPTI through torch.profiler; diagnostic only, never a bench value).

usage: python tools/trace_step.py [--workload C2] [--precision bf16] [--out gpurun_out/timeline.md]
Prints, for the median-length replay among those traced, every kernel with its stream, start offset and duration, and
the idle gaps of the device, so that the critical path of the step can be read off.
"""
import argparse
import json
import os
import re
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default="gpurun_out/timeline.md")
    ap.add_argument("--replays", type=int, default=7)
    ap.add_argument("--cells", type=int, default=None, help="cells per group held on the device (default: the workload's)")
    a = ap.parse_args()
    from spvipes_b200 import _lib as L
    from spvipes_b200 import synth
    from spvipes_b200.engine import GroupBatch, StepEngine
    from spvipes_b200.trainer import TrainLoop, init_params

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    mode, n_cells, genes, H, B, n_labels, _plan_dtype = bench.WORKLOADS[a.workload]
    n_cells = a.cells or n_cells
    L.load()
    data = synth.make_counts((n_cells, n_cells), (genes, genes), n_labels, device=dev, seed=1234)
    plan = None
    if mode != "label":  # OT modes: a plan over the cells held on the device, stored as the workload stores it
        plan = synth.make_plan(n_cells, n_cells, data.labels[0], data.labels[1], n_labels, device=dev, seed=7, dtype=_plan_dtype)
    eng = StepEngine((genes, genes), H, bench.S_DIM, bench.P_DIM, 0.1, mode, device=dev, seed=0, plan=plan, precision=a.precision)
    init_params(eng, 0)
    loop = TrainLoop(eng)
    loop.set_epoch(1)
    gen = torch.Generator(device=dev).manual_seed(5)
    rows_cur = [torch.randperm(n_cells, generator=gen, device=dev)[:B].to(torch.int32) for _ in (0, 1)]
    if mode == "label":
        batches = [GroupBatch(X=data.X[g], rows=rows_cur[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
    else:
        batches = [GroupBatch(X=data.X[g], rows=rows_cur[g], idx=rows_cur[g],
                              labels=data.labels[g][rows_cur[g].long()].contiguous() if mode == "cluster" else None) for g in (0, 1)]
    graph = loop.capture(batches)
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.replays):  # back to back, as in training: the host enqueues replay i+1 while replay i runs
            graph.replay()
        torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "t.json")
        prof.export_chrome_trace(path)
        tr = json.load(open(path))
    ks = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
    ks.sort(key=lambda e: e["ts"])
    n = len(ks) // a.replays  # every replay launches the same kernels
    groups = [ks[i * n:(i + 1) * n] for i in range(a.replays)][1:]
    spans = [g[-1]["ts"] + g[-1]["dur"] - g[0]["ts"] for g in groups]
    order = sorted(range(len(groups)), key=lambda i: spans[i])
    g = groups[order[len(order) // 2]]
    t0 = g[0]["ts"]
    streams = {}
    lines = [f"timeline of one replayed step ({a.workload}, {a.precision}): {len(g)} kernels, span "
             f"{spans[order[len(order) // 2]]:.1f} us (spans of all traced replays: {', '.join(f'{s:.0f}' for s in spans)})", "",
             "| start us | dur us | end us | stream | kernel | grid | block |", "|---:|---:|---:|---:|---|---|---|"]
    busy_end = 0.0
    idle = 0.0
    for e in g:
        s = e["args"].get("stream", 0)
        sid = streams.setdefault(s, len(streams))
        n = e["name"].replace("(anonymous namespace)::", "")
        n = re.sub(r"^void ", "", re.sub(r"\(.*", "", n))
        st, du = e["ts"] - t0, e["dur"]
        if st > busy_end:
            idle += st - busy_end
        busy_end = max(busy_end, st + du)
        lines.append(f"| {st:.1f} | {du:.1f} | {st + du:.1f} | {sid} | `{n}` | {e['args'].get('grid')} | {e['args'].get('block')} |")
    lines.insert(1, f"device fully idle for {idle:.1f} us of the span (no kernel of any stream running)")
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:3]))


if __name__ == "__main__":
    main()
