"""Benchmark of the contrastive-loss hot path (BASELINE.json metric: NT-Xent fwd+bwd views/s, 2N=8192, d=128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1  : one NT-Xent forward+backward over 2N = 8192 views, d = 128, tau = 0.5, fp32 inputs, bf16 tensor-core
         operands with fp32 accumulation.  `value` = views/s with inputs resident in HBM (the step is one
         CUDA graph of the six kernels, L2 flushed between steps, timed with CUDA events); `e2e` = the same
         step through the public `contrastive_loss` API from pinned HOST buffers, H2D copy and loss/accuracy
         read-back inside the timed region.
N > 1  : global batch 2N = 65536 sharded by rows over N ranks (torchrun, NCCL): all-gather of operands and
         lse2 inside the timed region, strong scaling ("scaling": "strong").
--impl reference : the reference algorithm on the host CPU (oracle dense port of objective.py, all cores).

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

TAU = 0.5
DIM = 128
B_SINGLE = 4096          # 2N = 8192  (BASELINE.json metric)
B_GLOBAL = 32768         # 2N = 65536 (BASELINE.json configs[3])
METRIC = "NT-Xent fwd+bwd views/s (2N=8192,d=128)"
L2_FLUSH_BYTES = 256 << 20


def algorithmic_flops(m, d):
    return 6.0 * m * m * d          # fwd 2M^2d + bwd row term 2M^2d + column term 2M^2d


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_arm(steps, warmup, b=B_SINGLE, d=DIM):
    """The reference's algorithm on the host cores: oracle/contrastive_oracle.py dense port (fp32, torch CPU)."""
    import torch
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import contrastive_oracle as oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(0)
    z1, z2 = torch.randn(b, d, generator=gen), torch.randn(b, d, generator=gen)

    def one():
        a = z1.clone().requires_grad_(True)
        c = z2.clone().requires_grad_(True)
        loss, acc = oracle.ntxent_dense_port(a, c, temperature=TAU)
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return 2 * b / dt, dt * 1e3, cores


def bench_single(args):
    import torch
    import pytorch_simclr_b200 as sb
    from pytorch_simclr_b200.functional import LOSS_NTXENT
    from pytorch_simclr_b200.runner import ContrastiveStep

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    b, d, m = B_SINGLE, DIM, 2 * B_SINGLE
    gen = torch.Generator().manual_seed(0)
    h1 = torch.randn(b, d, generator=gen).pin_memory()
    h2 = torch.randn(b, d, generator=gen).pin_memory()
    step = ContrastiveStep(LOSS_NTXENT, b, d, TAU, True, torch.float32, dev)
    step.x1.copy_(h1)
    step.x2.copy_(h2)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for _ in range(3):
            step.step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step.step()
    g_fwd, g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_fwd, stream=side):
        step.forward()
    with torch.cuda.graph(g_bwd, stream=side):
        step.backward()

    def timed(g, n, warm):
        for _ in range(warm):
            flush.zero_()
            g.replay()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        torch.cuda.synchronize()
        for i in range(n):
            flush.zero_()                   # evict the operands from L2 between steps
            starts[i].record()
            g.replay()
            stops[i].record()
        torch.cuda.synchronize()
        ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
        return sum(ms) / n, min(ms)

    sampler = ClockSampler(0)
    sampler.start()
    ms_step, ms_best = timed(graph, args.steps, args.warmup)
    ms_fwd, _ = timed(g_fwd, args.steps, 1)
    ms_bwd, _ = timed(g_bwd, args.steps, 1)

    # end to end through the public API: pinned host -> device, loss + accuracy read back
    def e2e_step():
        a = h1.to(dev, non_blocking=True).requires_grad_(True)
        c = h2.to(dev, non_blocking=True).requires_grad_(True)
        loss, acc = sb.contrastive_loss(a, c, temperature=TAU)
        loss.backward()
        return loss.item(), acc

    for _ in range(max(3, args.warmup)):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    # subtract nothing: the flush is part of the loop and is reported separately
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
    torch.cuda.synchronize()
    flush_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max(e2e_ms - flush_ms, 1e-6)
    clocks = sampler.stop()

    peak, peak_src = load_peaks()
    flops = algorithmic_flops(m, d)
    bwd_flops = 4.0 * m * m * d           # dominant kernel: backward tile kernel (row + column terms)
    achieved = bwd_flops / (ms_bwd * 1e-3) / 1e12
    cpu_value, cpu_ms, cores = cpu_reference_arm(steps=8, warmup=2)
    line = {
        "metric": METRIC, "value": m / (ms_step * 1e-3), "unit": "views/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ntxent_fwd_bwd 2N=8192 d=128 tau=0.5 fp32-in bf16-mma fp32-acc", "global_batch": b,
                   "l2": "flushed between steps (256 MiB memset outside the event pair)",
                   "launch": "one CUDA graph of 6 kernels per step"},
        "e2e": {"value": m / (e2e_ms * 1e-3), "unit": "views/s", "h2d_bytes_per_step": 2 * b * d * 4,
                "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms},
        "gpu_launches": 6 * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": None, "kernel": "contrastive_tile_kernel<128,0,true> (backward)",
                     "peak_source": peak_src, "ms_forward_stage": ms_fwd, "ms_backward_stage": ms_bwd,
                     "whole_step_tflops": flops / (ms_step * 1e-3) / 1e12,
                     "whole_step_frac": flops / (ms_step * 1e-3) / 1e12 / peak, "best_step_ms": ms_best},
        "cpu_baseline": {"value": cpu_value, "unit": "views/s", "cores": cores, "kind": "port",
                         "sample": "8 fwd+bwd calls of the same workload (2N=8192, d=128, fp32) after 2 warm-ups",
                         "ms_per_step": cpu_ms},
    }
    print(json.dumps(line))


def bench_multi(args):
    import torch
    import torch.distributed as dist
    from pytorch_simclr_b200 import functional as F
    from pytorch_simclr_b200.distributed import RowShardGather, global_contrastive_loss, shard_rows

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    b, d, m = B_GLOBAL, DIM, 2 * B_GLOBAL
    row_off, bl = shard_rows(b, world, rank)
    gen = torch.Generator().manual_seed(1000 + rank)
    h1 = torch.randn(bl, d, generator=gen).pin_memory()
    h2 = torch.randn(bl, d, generator=gen).pin_memory()
    x1, x2 = h1.to(dev), h2.to(dev)
    gather = RowShardGather()

    def step():
        loss, stats, _rv, saved = F.run_forward(F.LOSS_NTXENT, x1, x2, TAU, True, None, gather)
        return loss, F.run_backward(saved, x1, x2, None)

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([start.elapsed_time(stop) / args.steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t)

    def e2e_step():
        a = h1.to(dev, non_blocking=True).requires_grad_(True)
        c = h2.to(dev, non_blocking=True).requires_grad_(True)
        loss, acc = global_contrastive_loss(a, c, temperature=TAU)
        loss.backward()
        return loss.item()

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = load_peaks()
        flops = algorithmic_flops(m, d)
        achieved = flops / (ms_step * 1e-3) / 1e12 / world
        line = {
            "metric": METRIC, "value": m / (ms_step * 1e-3), "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ntxent_fwd_bwd global 2N=65536 d=128 tau=0.5 row-sharded", "global_batch": b,
                       "parallelism": f"rows/{world} + all-gather(operands, lse2)",
                       "l2": "operand matrix (16.8 MB) re-read per row block; inputs not flushed",
                       "collectives": "2x all_gather operands, 1x all_reduce stats, 2x all_gather lse2 (NCCL)"},
            "e2e": {"value": m / (e2e_ms * 1e-3), "unit": "views/s", "h2d_bytes_per_step": 2 * bl * d * 4 * world,
                    "d2h_bytes_per_step": 8 * world, "ms_per_step": e2e_ms},
            "gpu_launches": 6 * args.steps * world,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None, "kernel": "whole step per GPU (fwd+bwd tile kernels)",
                         "peak_source": peak_src},
        }
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 20)
        value, ms, cores = cpu_reference_arm(steps=steps, warmup=min(args.warmup, 3))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ntxent_fwd_bwd 2N=8192 d=128 tau=0.5 fp32 (reference algorithm, host CPU)",
                       "global_batch": B_SINGLE},
            "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} fwd+bwd calls of the workload on the host cores"},
            "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        bench_multi(args)
    else:
        bench_single(args)


if __name__ == "__main__":
    main()
