"""Achieved errors of both arithmetic modes against the fp64 oracle (test infrastructure: imports oracle/), and their
fwd+bwd time.  Output is committed as profiles/r01_precision.log."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import contrastive_oracle as oracle  # noqa: E402
import pytorch_simclr_b200 as sb  # noqa: E402
from pytorch_simclr_b200.functional import LOSS_MODIFIED, LOSS_NTXENT  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

cases = [("ntxent", 512, 128, 0.5, "iid"), ("ntxent", 4096, 128, 0.5, "iid"), ("ntxent", 4096, 128, 0.1, "correlated"),
         ("ntxent", 777, 100, 0.1, "correlated"), ("ntxent", 2048, 64, 0.5, "correlated"),
         ("modified", 4096, 128, 0.5, "iid"), ("modified", 4096, 128, 0.1, "correlated")]
print(f"{'loss':9s} {'B':>5s} {'d':>4s} {'tau':>4s} {'inputs':10s} {'mode':5s} | {'loss rel err':>12s} {'grad err / max|g|':>18s} {'acc diff (rows)':>15s} | us fwd+bwd")
for name, b, d, tau, kind in cases:
    z1, z2 = oracle.make_embeddings(b, d, seed=b + d, kind=kind)
    ref = (oracle.ntxent_closed_form if name == "ntxent" else oracle.modified_closed_form)(z1, z2, temperature=tau)
    fn = sb.contrastive_loss if name == "ntxent" else sb.modified_contrastive_loss
    for mode in ("bf16", "fp32"):
        sb.set_precision(mode)
        a = z1.cuda().requires_grad_(True)
        c = z2.cuda().requires_grad_(True)
        loss, acc = fn(a, c, temperature=tau)
        loss.backward()
        gmax = max(np.abs(ref.grad1).max(), np.abs(ref.grad2).max())
        e = max(np.abs(a.grad.cpu().numpy() - ref.grad1).max(), np.abs(c.grad.cpu().numpy() - ref.grad2).max()) / gmax
        lrel = abs(float(loss.detach()) - ref.loss) / abs(ref.loss)
        step = ContrastiveStep(LOSS_NTXENT if name == "ntxent" else LOSS_MODIFIED, b, d, tau, precision=mode)
        step.x1.copy_(z1)
        step.x2.copy_(z2)
        for _ in range(3):
            step.step()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(20):
            step.step()
        ev1.record()
        torch.cuda.synchronize()
        print(f"{name:9s} {b:5d} {d:4d} {tau:4.2f} {kind:10s} {mode:5s} | {lrel:12.2e} {e:18.2e} {abs(acc - ref.acc) * 2 * b / 100:15.1f} | "
              f"{ev0.elapsed_time(ev1) / 20 * 1e3:8.1f}")
sb.set_precision("auto")
