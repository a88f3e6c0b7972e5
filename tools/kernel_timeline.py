"""GPU-side timeline (ns, %globaltimer) of the kernels of one fwd+bwd step replayed as a CUDA graph,
with and without an L2 flush before the step.  Shows kernel durations AND the gaps between them."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the stamps only exist in the tracing build: make it the library the whole package uses in this process
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--loss", type=int, default=0)
args = ap.parse_args()
lib = _lib.load()
step = ContrastiveStep(args.loss, args.b, args.d, 0.5)
gen = torch.Generator().manual_seed(0)
step.x1.copy_(torch.randn(args.b, args.d, generator=gen))
step.x2.copy_(torch.randn(args.b, args.d, generator=gen))
names = ["prepare", "fwd_tile", "bwd_prepare", "bwd_tile", "fwd_fin", "bwd_fin"]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.simclr_debug_set_kernel_trace(buf.data_ptr())
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        step.step()
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    step.step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def reset():
    v = torch.zeros(16, dtype=torch.int64)
    v[0::2] = torch.iinfo(torch.int64).max
    buf.copy_(v)


for label, do_flush in (("warm L2", False), ("after L2 flush", True)):
    for rep in range(3):
        reset()
        if do_flush:
            flush.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        graph.replay()
        ev1.record()
        torch.cuda.synchronize()
        t = buf.cpu().view(8, 2)[:6]
        t0 = int(t[0, 0])
        line = " | ".join(f"{n} {int(a) - t0:6d}..{int(b) - t0:6d} ({int(b) - int(a):6d})" for n, (a, b) in zip(names, t.tolist())
                          if 0 < int(b) and int(a) < torch.iinfo(torch.int64).max)
        print(f"[{label}] events {ev0.elapsed_time(ev1) * 1e3:7.1f} us | {line}")
lib.simclr_debug_set_kernel_trace(None)
