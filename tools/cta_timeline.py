"""Per-CTA %globaltimer stamps of the forward / backward tile kernels (debug): where does the kernel's
duration come from -- tile work, fused finalize, or imbalance between CTAs?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the stamps are compiled out of the product library: use the tracing build (pytorch-simclr_b200/build.py makes both)
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

b, d = 4096, 128
lib = _lib.load()
step = ContrastiveStep(0, b, d, 0.5)
gen = torch.Generator().manual_seed(0)
step.x1.copy_(torch.randn(b, d, generator=gen))
step.x2.copy_(torch.randn(b, d, generator=gen))
for _ in range(3):
    step.step()
torch.cuda.synchronize()
buf = torch.zeros(64 + 148 * 8, dtype=torch.int64, device="cuda")
for phase in ("forward", "backward"):
    buf.zero_()
    buf[0:16:2] = torch.iinfo(torch.int64).max
    if phase == "backward":
        step.forward()
    torch.cuda.synchronize()
    lib.simclr_debug_set_kernel_trace(buf.data_ptr())
    getattr(step, phase)()
    torch.cuda.synchronize()
    lib.simclr_debug_set_kernel_trace(None)
    t = buf.cpu()[64:].view(148, 8).double()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    rel[t == 0] = float("nan")
    print(f"==== {phase}: per-CTA us relative to first CTA start")
    print("start  min/med/max: %.1f %.1f %.1f" % (rel[:, 0].min(), rel[:, 0].median(), rel[:, 0].max()))
    for k, name in ((1, "seg0 tiles done"), (2, "seg0 finalize done"), (3, "seg1 tiles done"), (4, "seg1 finalize done"), (5, "end")):
        col = rel[:, k]
        col = col[~col.isnan()]
        if len(col):
            print(f"{name:20s} n={len(col):3d} min/med/max: {col.min():.1f} {col.median():.1f} {col.max():.1f}")
    fin0 = (rel[:, 2] - rel[:, 1])
    fin1 = (rel[:, 4] - rel[:, 3])
    for name, f in (("seg0 finalize dur", fin0), ("seg1 finalize dur", fin1)):
        f = f[~f.isnan()]
        print(f"{name:20s} min/med/max: {f.min():.2f} {f.median():.2f} {f.max():.2f}  (>1us: {(f > 1).sum().item()})")
    raw = buf.cpu()[64:].view(148, 8)
    mhz = (raw[:, 7] - raw[:, 6]).double() / (raw[:, 5] - raw[:, 0]).double() * 1e3
    print("SM clock during the kernel (clock64 / globaltimer): min/med/max MHz %.0f %.0f %.0f" % (mhz.min(), mhz.median(), mhz.max()))
    worst = torch.argsort(rel[:, 5], descending=True)[:5]
    for c in worst.tolist():
        print(f"  slow CTA {c}: " + " ".join(f"{x:.1f}" for x in rel[c, :6].tolist()))
