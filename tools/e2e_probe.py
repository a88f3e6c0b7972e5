"""Host-side cost of the public API path (autograd Function -> ctypes -> C ABI) at a size where the GPU work is
negligible: where do the microseconds of `e2e` go?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pytorch_simclr_b200 as sb  # noqa: E402
from pytorch_simclr_b200 import functional as F  # noqa: E402

sb.set_precision("bf16")
dev = torch.device("cuda", 0)
b = 128
a = torch.randn(b, 128, device=dev, requires_grad=True)
c = torch.randn(b, 128, device=dev, requires_grad=True)
N = 300


def timeit(name, fn):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    torch.cuda.synchronize()
    print(f"{name:48s} {(time.perf_counter() - t0) / N * 1e6:8.1f} us")


saved_box = {}


def fwd_raw():
    loss, stats, rv, saved = F.run_forward(F.LOSS_NTXENT, a, c, 0.5, True, None, None, True)
    saved_box["s"] = saved
    return loss


def fwd_bwd_raw():
    fwd_raw()
    F.run_backward(saved_box["s"], a, c, None)


def fn_apply():
    loss, stats = F.ContrastiveLossFunction.apply(a, c, F.LOSS_NTXENT, 0.5, True, None, None)
    return loss


def fn_apply_bwd():
    fn_apply().backward()


timeit("run_forward (no autograd, no sync)", fwd_raw)
timeit("run_forward + run_backward (no autograd)", fwd_bwd_raw)
timeit("Function.apply forward", fn_apply)
timeit("Function.apply + loss.backward()", fn_apply_bwd)
timeit("contrastive_loss (forward + acc .item())", lambda: sb.contrastive_loss(a, c, temperature=0.5))
timeit("contrastive_loss + backward + loss.item()", lambda: (lambda l: (l[0].backward(), l[0].item()))(sb.contrastive_loss(a, c, temperature=0.5)))
x = torch.randn(8, device=dev, requires_grad=True)
timeit("reference point: (x*x).sum().backward()", lambda: (x * x).sum().backward())
