"""Kernel timeline (tracing build) of the fused step for chosen input sets of tools/per_set_times.py: which kernel pays for a
row whose accuracy decision needs the exact re-scoring?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

lib = _lib.load()
b, d = 4096, 128
step = ContrastiveStep(0, b, d, 0.5)
dgen = torch.Generator(device="cuda").manual_seed(1)
sets = [(torch.randn(b, d, generator=dgen, device="cuda"), torch.randn(b, d, generator=dgen, device="cuda"),
         torch.empty(b, d, device="cuda"), torch.empty(b, d, device="cuda")) for _ in range(8)]
names = ["prepare", "fwd_tile", "bwd_prepare", "bwd_tile", "fwd_fin", "bwd_fin"]
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
hdr = step.fwd_ws[:16].view(torch.int32)
lib.simclr_debug_set_kernel_trace(buf.data_ptr())
for i in (0, 2, 4, 5):
    x1, x2, g1, g2 = sets[i]
    for rep in range(3):
        v = torch.zeros(2048, dtype=torch.int64)
        v[0:16:2] = torch.iinfo(torch.int64).max
        buf.copy_(v)
        flush.zero_()
        torch.cuda.synchronize()
        step.step(None, x1, x2, g1, g2)
        torch.cuda.synchronize()
    t = buf.cpu()[:16].view(8, 2)[:6]
    t0 = int(t[0, 0])
    line = " | ".join(f"{n} {int(a) - t0:6d}..{int(z) - t0:6d} ({int(z) - int(a):6d})" for n, (a, z) in zip(names, t.tolist())
                      if 0 < int(z) and int(a) < torch.iinfo(torch.int64).max)
    print(f"set {i}: listed {hdr.cpu().tolist()[1]} | {line}")
lib.simclr_debug_set_kernel_trace(None)
