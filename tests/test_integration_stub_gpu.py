"""The ctypes stub printed in INTEGRATION.md section 2 is executed as written (only the library path is filled in) and must
reproduce the package's own result: the document cannot drift from the ABI."""
import os
import re

import pytest
import torch

import pytorch_simclr_b200 as sb
from pytorch_simclr_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source() -> str:
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "objective_b200.py" in b]
    assert len(stub) == 1
    return stub[0].replace("/path/to/libsimclr_b200.so", _lib.LIB_PATH)


def test_integration_stub_runs_and_matches_the_package():
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md:stub", "exec"), ns)
    g = torch.Generator().manual_seed(7)
    b, d = 300, 128
    x1 = torch.randn(b, d, generator=g).cuda()
    x2 = (x1.cpu() + 0.7 * torch.randn(b, d, generator=g)).cuda()
    a1, a2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    loss_s, acc_s = ns["contrastive_loss"](a1, a2, temperature=0.5)
    (loss_s / 4).backward()
    sb.set_precision("bf16")
    try:
        c1, c2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
        loss_p, acc_p = sb.contrastive_loss(c1, c2, temperature=0.5)
        (loss_p / 4).backward()
    finally:
        sb.set_precision("auto")
    torch.cuda.synchronize()
    assert torch.equal(loss_s, loss_p.detach())
    assert acc_s == acc_p
    # (the default backward adds the accumulators of the CTAs that share a row block in the order they finish: last-bit noise)
    scale = float(c1.grad.abs().max())
    assert float((a1.grad - c1.grad).abs().max()) <= 1e-5 * scale
    assert float((a2.grad - c2.grad).abs().max()) <= 1e-5 * scale
