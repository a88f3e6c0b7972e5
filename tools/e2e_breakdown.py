"""Where the microseconds of `e2e` go (bench.py's end-to-end arm): the public API from pinned host buffers, phase by
phase on the host clock, for both arithmetic modes, alternating (so that an order / warm-up effect shows), optionally with
the nvidia-smi clock sampler bench.py runs next to it.

    python tools/e2e_breakdown.py [--sampler]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pytorch_simclr_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sampler", action="store_true")
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()
dev = torch.device("cuda", 0)
b, d = 4096, 128
g = torch.Generator().manual_seed(0)
h12 = torch.stack((torch.randn(b, d, generator=g), torch.randn(b, d, generator=g))).pin_memory()
sampler = None
if args.sampler:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ClockSampler
    sampler = ClockSampler(0)
    sampler.start()


def run(mode, n):
    sb.set_precision(mode)
    marks = [0.0] * 6
    gpu = [0.0] * 3
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def one(record):
        t0 = time.perf_counter()
        ev[0].record()
        x = h12.to(dev, non_blocking=True)
        ev[1].record()
        a = x[0].requires_grad_(True)
        c = x[1].requires_grad_(True)
        t1 = time.perf_counter()
        loss, acc = sb.contrastive_loss(a, c, temperature=0.5)
        ev[2].record()
        t2 = time.perf_counter()
        loss.backward()
        ev[3].record()
        t3 = time.perf_counter()
        v = loss.item()
        t4 = time.perf_counter()
        if record:
            for i, dt in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                marks[i] += dt
            for i in range(3):
                gpu[i] += ev[i].elapsed_time(ev[i + 1]) * 1e-3
        return v

    for _ in range(10):
        one(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        one(True)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / n * 1e6
    names = ("h2d+views", "contrastive_loss (incl. acc read-back)", "loss.backward()", "loss.item()")
    print(f"[{mode:5s}] {total:7.1f} us/step | " + " | ".join(f"{nm} {marks[i] / n * 1e6:6.1f}" for i, nm in enumerate(names))
          + " || GPU stream: h2d {:.1f} | fwd call (4 kernels) {:.1f} | bwd call (1 kernel) {:.1f}".format(*(g / n * 1e6 for g in gpu)),
          flush=True)


for rep in range(3):
    for mode in ("bf16", "fp32", "auto"):
        run(mode, args.steps)
if sampler is not None:
    print(sampler.stop())
