#!/usr/bin/env python
"""Benchmark of the spVIPES per-minibatch training hot path on B200 (BASELINE.json metric: training cells/sec, fwd+bwd
ELBO + optimiser; NB-loglik kernel HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

A "step" is one minibatch (B cells per group, 2 groups) through inference -> generative -> loss -> backward -> Adam.
Workload (N = 1): BASELINE.json configs[1] — label-based PoE, 2 groups x 50k cells, 5k HVGs, n_hidden 128, batch 512.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mode, cells per group, genes per group, n_hidden, batch per group, n_labels)
    "C1": ("label", 5_000, 2_000, 128, 512, 10),
    "C2": ("label", 50_000, 5_000, 128, 512, 10),
    "C5": ("label", 1_000_000, 20_000, 128, 2048, 10),
}
S_DIM, P_DIM = 25, 10
METRIC = "training cells/sec (fwd+bwd ELBO + Adam)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, through NVML in a background thread (pynvml, 20 ms
    period).  A resident `nvidia-smi -lms` poller was measured to stall the CUDA launch path of the benchmarked process
    (2-GPU step 0.46 ms without it, 0.60 ms with it at 20 ms, 1.28 ms at 100 ms), so it is only the fallback."""

    def __init__(self, index=0):
        self.index, self.samples, self.reason_bits, self.max_mhz = index, [], 0, None
        self.thread, self.stop_flag, self.proc, self.rows = None, False, None, []

    def start(self):
        if os.environ.get("SPV_NO_CLOCKS"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self.stop_flag:
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.reason_bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.02)

            self._nvml = pynvml
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # the first line arrives ~100 ms after start
                time.sleep(0.01)
            self.n_before = len(self.rows)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            nv = self._nvml
            names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": [n for n, bit in names if self.reason_bits & int(bit)], "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.rows = self.rows[getattr(self, "n_before", 0):] or self.rows[-1:]
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_cells_per_sec(workload, steps, warmup, seed=0):
    """oracle/restatement.py (validated against the unmodified reference, tests/test_oracle.py) + autograd backward + Adam
    (scvi TrainingPlan defaults) with every host thread; bounded sample: `steps` minibatches of the workload's shape."""
    from oracle import restatement as rs
    from spvipes_b200 import synth
    from spvipes_b200.engine import StepEngine
    from spvipes_b200.trainer import init_params

    mode, n_cells, genes, H, B, n_labels = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_sample = B * (steps + warmup)
    data = synth.make_counts((n_sample, n_sample), (genes, genes), n_labels, device="cpu", seed=1234)
    eng = StepEngine((genes, genes), H, S_DIM, P_DIM, 0.1, mode, device="cpu")
    init_params(eng, seed)
    sd = {k: v.clone() for k, v in eng.state_dict().items()}
    names = rs.param_names(sd)
    for k in names:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in names], lr=1e-3, eps=0.01, weight_decay=1e-6)
    gen = torch.Generator().manual_seed(1)
    t0 = None
    for s in range(steps + warmup):
        if s == warmup:
            t0 = time.perf_counter()
        sl = slice(s * B, (s + 1) * B)
        x = [data.X[g][sl].to(torch.float32) for g in (0, 1)]
        labels = [data.labels[g][sl].numpy() for g in (0, 1)]
        eps_p = [torch.randn(B, P_DIM, generator=gen) for _ in (0, 1)]
        eps_q = [torch.randn(B, S_DIM, generator=gen) for _ in (0, 1)]
        dm = {(g, k): (torch.rand(B, H, generator=gen) < 0.9).float() / 0.9 for g in (0, 1) for k in ("private", "shared")}
        out = rs.step(sd, x, mode=mode, n_shared=S_DIM, n_private=P_DIM, eps_private=eps_p, eps_poe=eps_q, labels=labels,
                      drop_masks=dm, kl_weight=min(1.0, s / 400.0))
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()
        for k, v in out["new_stats"].items():
            sd[k] = v
    dt = time.perf_counter() - t0
    return 2 * B * steps / dt, dt / steps * 1e3, cores, f"{steps} minibatches of {workload} shape (2x{B} cells, {genes} genes/group), oracle port, torch CPU"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    steps = min(args.steps, 20)
    v, ms, cores, sample = cpu_reference_cells_per_sec(workload, steps, min(args.warmup, 2))
    mode, n_cells, genes, H, B, n_labels = WORKLOADS[workload]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(workload, args.gpus),
            "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(workload, n_gpus):
    mode, n_cells, genes, H, B, n_labels = WORKLOADS[workload]
    return {"workload": f"{workload}: {mode}-based PoE, 2 groups x {n_cells} cells, {genes} genes/group, n_hidden {H}, "
                        f"shared {S_DIM} / private {P_DIM}, batch {B}/group/GPU, {n_labels} labels",
            "cells_per_step_per_gpu": 2 * B, "parallelism": f"dp{n_gpus}",
            "l2": "inputs larger than L2: every step gathers a fresh random minibatch from the device-resident count "
                  "matrices and streams all weights/Adam state"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from spvipes_b200 import _lib as L
    from spvipes_b200 import synth
    from spvipes_b200.engine import GroupBatch, StepEngine
    from spvipes_b200.trainer import TrainLoop, init_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("SPV_ALL_ON_GPU0"):
        local_rank = 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        backend = os.environ.get("SPV_DIST_BACKEND", "nccl")  # "gloo": debugging the multi-rank flow on a single GPU
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group(backend)
    workload = args.workload
    mode, n_cells, genes, H, B, n_labels = WORKLOADS[workload]
    if world > 1 and workload == "C5":
        n_cells = n_cells // world  # rank-local shard of the 2 x 1M cells (weak scaling: batch per GPU fixed)
    lib = L.load()
    L.check(lib.spv_arch_check(local_rank), "spv_arch_check (this library is sm_100a only)")
    data = synth.make_counts((n_cells, n_cells), (genes, genes), n_labels, device=dev, seed=1234 + 17 * rank)
    eng = StepEngine((genes, genes), H, S_DIM, P_DIM, 0.1, mode, device=dev, seed=rank, precision=args.precision)
    init_params(eng, 0)
    loop = TrainLoop(eng)
    if world > 1:
        from spvipes_b200.parallel import GradSync
        loop.grad_sync = GradSync(eng, dist)
    K, W = args.steps, args.warmup
    total = K + W
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    rows = [torch.stack([torch.randperm(n_cells, generator=gen, device=dev)[:B] for _ in range(total)]).to(torch.int32)
            for _ in (0, 1)]

    def batches_for(s):
        return [GroupBatch(X=data.X[g], rows=rows[g][s], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    loop.set_epoch(1)
    use_graph = not args.no_graph
    rows_cur = [torch.empty(B, dtype=torch.int32, device=dev) for _ in (0, 1)]
    static_batches = [GroupBatch(X=data.X[g], rows=rows_cur[g], labels=data.labels[g], labels_per_cell=True) for g in (0, 1)]
    per_step_launches = None
    if use_graph:
        for g in (0, 1):
            rows_cur[g].copy_(rows[g][0])
        c0 = lib.spv_launch_count()
        graph = loop.capture(static_batches)
        # 2 warm-up steps + 1 captured step (+ the 4 bf16 weight re-staging launches at the end of capture())
        per_step_launches = (lib.spv_launch_count() - c0 - (4 if eng.bf16 else 0)) // 3

    def run_step(s):
        if use_graph:
            for g in (0, 1):
                rows_cur[g].copy_(rows[g][s], non_blocking=True)
            graph.replay()
        else:
            loop.step(batches_for(s))

    for s in range(W):
        run_step(s)
    barrier()
    clk = ClockSampler(local_rank)
    if rank == 0:
        clk.start()
    n0 = lib.spv_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(W, total):
        run_step(s)
    ev1.record()
    barrier()
    launches = per_step_launches * K if use_graph else lib.spv_launch_count() - n0
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = clk.stop() if rank == 0 else None
    loss = float(eng.loss_out[0].item())
    value = world * 2 * B * K / (ms * 1e-3)

    # ---- the NB-loglik kernel alone, timed with CUDA events on the launching stream: K eager forward passes over fresh
    #      minibatches (the graph replays above cannot carry timing events); every launch of the sweep is bracketed.
    nb_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * K)]
    eng.nb_events = iter(nb_ev)
    eng.parallel_groups = False  # the kernel is timed alone: no second group running beside it
    for s in range(W, total):
        eng.forward(batches_for(s), training=True)
    torch.cuda.synchronize()
    eng.nb_events = None
    eng.parallel_groups = True
    # ---- NB-loglik kernel roofline (forward sweep of the fused decoder + NB kernel), timed live with CUDA events
    nb_ms = float(np.mean([a.elapsed_time(b) for a, b in nb_ev]))
    KM = 256 + S_DIM + P_DIM
    # algorithmic bytes of one launch: counts u16 + mixture logits written for the backward (f32) + weights
    # (mixture weight + folded factor-regressor weights) + per-gene constants + decoder inputs + per-row outputs
    if args.precision == "bf16":
        KMp = (KM + 7) // 8 * 8
        # counts u16 + stacked bf16 weights (mixture + two folded branch blocks) + per-gene constants + [hm | zz] bf16 + rows
        alg_bytes = B * genes * 2 + 3 * genes * KMp * 2 + 6 * genes * 4 + B * KMp * 2 + B * 12 * 4
        kname = "nb_tc_fwd_kernel (tcgen05 decoder GEMMs + fused NB-mixture log-likelihood epilogue, forward)"
    else:
        alg_bytes = B * genes * 2 + B * genes * 4 + genes * KM * 4 + genes * 35 * 4 + 6 * genes * 4 + B * KM * 4 + B * 12 * 4
        kname = "dec_tile_kernel<PASS_NB> (fp32 SIMT decoder GEMM + fused NB-mixture log-likelihood, forward)"
    peak, peak_src = peaks()
    achieved = alg_bytes / (nb_ms * 1e-3) / 1e9
    # dram__bytes_read + write per launch from the committed ncu --set full capture of this workload (profiles/r1_nb_final_ncu.md)
    traffic = 11.05e6 if (workload == "C2" and args.precision == "bf16") else None
    roofline = {"kernel": kname, "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": nb_ms,
                "note": "instruction-issue bound, not HBM bound: 132 warp instructions per 32 (cell, gene) elements, 16 of them MUFU "
                        "at 8 issue cycles each (tools/ubench/pipes.cu, profiles/r1_nb_persistent_notes.md, "
                        "profiles/r1_nb_final_ncu.md); the HBM fraction is reported as the contract asks"}
    # ---- the step's HBM-bound kernel for comparison: Adam over the whole flat parameter vector (28 bytes per parameter),
    #      timed alone with CUDA events (lr = 0: the parameters stay put, the traffic is the same)
    ad_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in ad_ev:
        a.record()
        eng.adam_step(lr=0.0, eps=0.01, weight_decay=0.0)
        b.record()
    torch.cuda.synchronize()
    ad_ms = float(np.median([a.elapsed_time(b) for a, b in ad_ev[2:]]))
    ad_bytes = 28 * eng.params.numel + (2 * eng.params.numel if args.precision == "bf16" else 0)
    roofline["other_kernels"] = [{"kernel": "adam_kernel (whole parameter vector, + bf16 operand staging)", "bound": "hbm",
                                  "achieved": ad_bytes / (ad_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": ad_bytes / (ad_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": ad_bytes,
                                  "avg_launch_ms": ad_ms}]

    # ---- end-to-end through the public step API with HOST (pinned) minibatches
    e2e = None if args.no_e2e else measure_e2e(loop, data, rows, B, genes, K, W, dev, dist, world, use_graph)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cms, cores, sample = cpu_reference_cells_per_sec(workload, args.cpu_steps, 1)
            cpu = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": cms}
        line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 tensor-core GEMMs, f32 accumulate / elementwise / Adam" if args.precision == "bf16" else "f32",
                "data": "synthetic", "config": workload_config(workload, world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "final_loss": loss}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure_e2e(loop, data, rows, B, genes, K, W, dev, dist, world, use_graph=True):
    """same metric through the public step call with host buffers: per step the two groups' count minibatches and labels are
    copied from pinned host memory (double-buffered on a copy stream) and the loss terms are read back."""
    from spvipes_b200.engine import GroupBatch
    total = K + W
    nh = min(total, 32)  # distinct pinned host minibatches, cycled (every step still copies its minibatch host -> device)
    host_x = [[data.X[g].view(torch.int16)[rows[g][s].long()].view(torch.uint16).cpu().pin_memory() for s in range(nh)]
              for g in (0, 1)]
    host_l = [[data.labels[g][rows[g][s].long()].cpu().pin_memory() for s in range(nh)] for g in (0, 1)]
    dev_x = [[torch.empty(B, genes, dtype=torch.uint16, device=dev) for _ in (0, 1)] for _ in (0, 1)]  # [buf][group]
    dev_l = [[torch.empty(B, dtype=torch.int32, device=dev) for _ in (0, 1)] for _ in (0, 1)]
    out_host = torch.empty(8, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in (0, 1)]
    freed = [torch.cuda.Event() for _ in (0, 1)]
    main = torch.cuda.current_stream(dev)
    bufs = [[GroupBatch(X=dev_x[b][g], labels=dev_l[b][g]) for g in (0, 1)] for b in (0, 1)]
    graphs = None
    if use_graph:
        for b in (0, 1):
            for g in (0, 1):
                dev_x[b][g].copy_(host_x[g][0]); dev_l[b][g].copy_(host_l[g][0])
        graphs = [loop.capture(bufs[b]) for b in (0, 1)]

    def upload(s):
        b = s % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            for g in (0, 1):
                dev_x[b][g].copy_(host_x[g][s % nh], non_blocking=True)
                dev_l[b][g].copy_(host_l[g][s % nh], non_blocking=True)
            ready[b].record(copy_stream)

    for b in (0, 1):
        freed[b].record(main)
    h2d = sum(host_x[g][0].numel() * 2 + host_l[g][0].numel() * 4 for g in (0, 1))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    upload(0)
    for s in range(total):
        if s == W:
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            ev0.record(main)
            upload(s)  # the first timed step's copy happens inside the timed region
        b = s % 2
        if s + 1 < total and s + 1 != W:
            upload(s + 1)
        main.wait_event(ready[b])
        if graphs is not None:
            graphs[b].replay()
        else:
            loop.step(bufs[b])
        freed[b].record(main)
        out_host.copy_(loop.engine.loss_out, non_blocking=True)
    ev1.record(main)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * 2 * B * K / (ms * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 32, "ms_per_step": ms / K,
            "api": "spvipes_b200.trainer.TrainLoop (host uint16 minibatches in pinned memory, loss terms read back)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying a CUDA graph")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: tcgen05 tensor-core path for the large GEMMs (parity gate 1e-2); fp32: SIMT path (1e-4)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-fed measurement")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "C2"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
