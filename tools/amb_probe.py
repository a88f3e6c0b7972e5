"""How many rows does the fused step hand to the exact re-scoring, and what does it cost the backward tile kernel?
Reads the forward-workspace header (word 1 = rows listed, word 2 = rows confirmed) after one fused step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

b, d = 4096, 128
for kind in (1, 0):
    for data in ("iid", "correlated"):
        step = ContrastiveStep(kind, b, d, 0.5, True, torch.float32, "cuda", "bf16")
        g = torch.Generator().manual_seed(b + d)
        if data == "iid":
            step.x1.copy_(torch.randn(b, d, generator=g))
            step.x2.copy_(torch.randn(b, d, generator=g))
        else:
            base = torch.randn(b, d, generator=g)
            step.x1.copy_(base + 0.5 * torch.randn(b, d, generator=g))
            step.x2.copy_(base + 0.5 * torch.randn(b, d, generator=g))
        for _ in range(3):
            step.step()
        torch.cuda.synchronize()
        hdr = step.fwd_ws[:16].view(torch.int32).cpu().tolist()
        bp = (b + 127) // 128 * 128
        # candidate counters as the forward tile kernel leaves them (prepare zeroes them; staged calls)
        step.prepare()
        step.forward(_lib.STAGE_FORWARD_TILE)
        torch.cuda.synchronize()
        off = 256 + (2 * bp // 128 * 16 + 255) // 256 * 256
        cand_cnt = step.fwd_ws[off:off + 2 * bp * 4].view(torch.int32).cpu()
        hist = torch.bincount(cand_cnt.clamp(max=9), minlength=10).tolist()
        step.step()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            step.step()
        ev[1].record()
        torch.cuda.synchronize()
        print(f"loss {kind} {data:10s}: header (ticket, rows listed, -, -) {hdr} | rows by published candidate chunks (0..8, 9+) {hist} | "
              f"correct {float(step.stats[2]):.0f} | {ev[0].elapsed_time(ev[1]) * 1e3 / 20:.1f} us per eager step", flush=True)
