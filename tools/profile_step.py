"""Tiny driver for ncu: a few NT-Xent (and optionally modified-loss) fwd+bwd steps at the BASELINE shape.

    python tools/profile_step.py [--b 4096] [--d 128] [--steps 3] [--loss 0]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--loss", type=int, default=0)
ap.add_argument("--tau", type=float, default=0.5)
args = ap.parse_args()

step = ContrastiveStep(args.loss, args.b, args.d, args.tau)
gen = torch.Generator().manual_seed(0)
step.x1.copy_(torch.randn(args.b, args.d, generator=gen))
step.x2.copy_(torch.randn(args.b, args.d, generator=gen))
for _ in range(args.steps):
    step.step()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(args.steps):
    step.step()
ev1.record()
torch.cuda.synchronize()
print(f"loss {float(step.loss):.6f} correct {float(step.stats[2]):.0f}  {ev0.elapsed_time(ev1) / args.steps * 1e3:.1f} us/step (eager launches)")
