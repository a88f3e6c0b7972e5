"""Print the per-role clock64() timeline of one CTA of the forward and backward tile kernels.

    python tools/trace_timeline.py [--b 4096] [--d 128] [--cta 5] [--loss 0]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the stamps are compiled out of the product library: use the tracing build (pytorch-simclr_b200/build.py makes both)
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--cta", type=int, default=5)
ap.add_argument("--loss", type=int, default=0)
args = ap.parse_args()
lib = _lib.load()
step = ContrastiveStep(args.loss, args.b, args.d, 0.5)
gen = torch.Generator().manual_seed(0)
step.x1.copy_(torch.randn(args.b, args.d, generator=gen))
step.x2.copy_(torch.randn(args.b, args.d, generator=gen))
for _ in range(3):
    step.step()
torch.cuda.synchronize()
ROLES, ITERS = 6, 64
names = ["tma", "mma", "wg0", "wg1", "wg2", "wg3"]
for phase in ("forward", "backward"):
    buf = torch.zeros(ROLES * ITERS * 4, dtype=torch.int64, device="cuda")
    if phase == "backward":
        step.forward()
        torch.cuda.synchronize()
    lib.simclr_debug_set_trace(buf.data_ptr(), args.cta)
    getattr(step, phase)()
    torch.cuda.synchronize()
    lib.simclr_debug_set_trace(None, 0)
    t = buf.cpu().view(ROLES, ITERS, 4)
    nz = t[t > 0]
    t0 = int(nz.min())
    print(f"==== {phase}: CTA {args.cta}, cycles relative to first event; span {int(nz.max()) - t0} cycles")
    print("it | tma: wait_b_empty issue | mma: enter issue [wait_w gradissue] | wgX: wait_s got_s done token")
    for it in range(ITERS):
        if not (t[:, it] > 0).any():
            break
        row = [f"{it:2d}"]
        for r in range(ROLES):
            vals = [int(v) - t0 if v > 0 else None for v in t[r, it]]
            if any(v is not None for v in vals):
                row.append(names[r] + ":" + ",".join("-" if v is None else str(v) for v in vals))
        print(" | ".join(row))
