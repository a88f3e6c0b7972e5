"""Diagnostics (tracing build): how often does the forward tile kernel's candidate recording leave its hot path?
Counts per forward launch: first-level triggers (a tile maximum reaches an undecided row's band or exceeds it), rescans
(a tile maximum INSIDE a band: the warp re-reads its 64 columns) and rows found in band."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

lib = _lib.load()
for kind, label in ((0, "ntxent"), (1, "modified")):
    for data in ("iid", "correlated"):
        b, d = 4096, 128
        step = ContrastiveStep(kind, b, d, 0.5)
        g = torch.Generator().manual_seed(0)
        if data == "iid":
            step.x1.copy_(torch.randn(b, d, generator=g))
            step.x2.copy_(torch.randn(b, d, generator=g))
        else:
            base = torch.randn(b, d, generator=g)
            step.x1.copy_(base + 0.5 * torch.randn(b, d, generator=g))
            step.x2.copy_(base + 0.5 * torch.randn(b, d, generator=g))
        buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
        v = torch.zeros(2048, dtype=torch.int64)
        v[0:16:2] = torch.iinfo(torch.int64).max
        buf.copy_(v)
        traced = hasattr(lib, "simclr_debug_set_kernel_trace")
        if traced:
            lib.simclr_debug_set_kernel_trace(buf.data_ptr())
        step.step()
        torch.cuda.synchronize()
        if traced:
            lib.simclr_debug_set_kernel_trace(None)
        t = buf.cpu()
        # the forward tile kernel alone, 20 launches between two events (no tracing)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            e0.record()
            for _ in range(20):
                step.forward(_lib.STAGE_FORWARD_TILE)
            e1.record()
            torch.cuda.synchronize()
        alone_us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"{label:8s} {data:10s}: first-level triggers {int(t[32]):8d}  rescans {int(t[33]):8d}  rows in band {int(t[34]):8d}  cycles in rescans {int(t[35]):9d} (inside routine {int(t[36])}, in {int(t[38])} TMEM loads {int(t[37])})  "
              f"(warp-tiles: {64 * 64 * 16 // (1 if kind == 0 else 2)}; fwd tile {int(t[3]) - int(t[2])} ns, alone {alone_us:.1f} us) acc rows {float(step.stats[2]):.0f}")
