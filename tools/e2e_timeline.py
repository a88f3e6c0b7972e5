"""GPU-side timeline (ns, %globaltimer, tracing build) of the kernels of ONE end-to-end step through the public API
(pinned host -> device, contrastive_loss, backward, loss.item()): when does each kernel start relative to the end of the
H2D copy, how long does it run, where are the gaps?

    python tools/e2e_timeline.py [--precision bf16|fp32]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402

import pytorch_simclr_b200 as sb  # noqa: E402
from pytorch_simclr_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
lib = _lib.load()
sb.set_precision(args.precision)
dev = torch.device("cuda", 0)
b, d = 4096, 128
g = torch.Generator().manual_seed(0)
h12 = torch.stack((torch.randn(b, d, generator=g), torch.randn(b, d, generator=g))).pin_memory()
names = ["prepare", "fwd_tile", "bwd_prepare", "bwd_tile", "fwd_fin", "bwd_fin"]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
stamp = torch.zeros(2, dtype=torch.int64, device="cuda")       # %globaltimer written by a tiny torch op is not available:
lib.simclr_debug_set_kernel_trace(buf.data_ptr())              # the copy's end is bracketed with events instead


def reset():
    v = torch.zeros(16, dtype=torch.int64)
    v[0::2] = torch.iinfo(torch.int64).max
    buf.copy_(v)


def step(record):
    t0 = time.perf_counter()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    x = h12.to(dev, non_blocking=True)
    e1.record()
    a = x[0].requires_grad_(True)
    c = x[1].requires_grad_(True)
    loss, acc = sb.contrastive_loss(a, c, temperature=0.5)
    t1 = time.perf_counter()
    loss.backward()
    e2.record()
    t2 = time.perf_counter()
    v = loss.item()
    t3 = time.perf_counter()
    if record:
        torch.cuda.synchronize()
        t = buf.cpu().view(8, 2)[:6]
        start = min(int(a) for a, z in t.tolist() if 0 < int(z))
        line = " | ".join(f"{n} +{int(a) - start:6d} ({int(z) - int(a):6d})" for n, (a, z) in zip(names, t.tolist())
                          if 0 < int(z) and int(a) < torch.iinfo(torch.int64).max)
        print(f"host: loss returned {1e6 * (t1 - t0):6.1f} backward returned {1e6 * (t2 - t0):6.1f} item {1e6 * (t3 - t0):6.1f} us | "
              f"events: h2d {e0.elapsed_time(e1) * 1e3:6.1f} us, h2d end -> bwd_fin end {e1.elapsed_time(e2) * 1e3:6.1f} us | {line}")


for _ in range(20):
    step(False)
for _ in range(6):
    torch.cuda.synchronize()
    reset()
    torch.cuda.synchronize()
    step(True)
lib.simclr_debug_set_kernel_trace(None)
