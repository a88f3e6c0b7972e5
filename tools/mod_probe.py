"""Per-kernel times of the modified loss (2N=8192, d=128) by input dtype and temperature: fused step (graph, L2 flushed),
and every kernel of the step launched alone 20x (stage masks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

b, d = 4096, 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()


def alone(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


for kind in (1, 0):
    for dtype in (torch.float32, torch.bfloat16):
        for tau in (0.5,):
            step = ContrastiveStep(kind, b, d, tau, True, dtype, "cuda", "bf16")
            g = torch.Generator().manual_seed(b + d)
            step.x1.copy_(torch.randn(b, d, generator=g))
            step.x2.copy_(torch.randn(b, d, generator=g))
            with torch.cuda.stream(side):
                for _ in range(3):
                    step.step()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                step.step()
            ts = []
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            # the same 20 replays queued without a host synchronisation in between (bench.py's protocol)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for a, c in evs:
                flush.zero_()
                a.record()
                gr.replay()
                c.record()
            torch.cuda.synchronize()
            tq = [a.elapsed_time(c) * 1e3 for a, c in evs]
            print("   synchronised:", " ".join(f"{t:.0f}" for t in ts), "| queued:", " ".join(f"{t:.0f}" for t in tq))
            ts.sort()
            step.step_staged()
            torch.cuda.synchronize()
            parts = {
                "prepare": alone(step.prepare),
                "fwd_tile": alone(lambda: step.forward(_lib.STAGE_FORWARD_TILE)),
                "fwd_fin": alone(lambda: step.forward(_lib.STAGE_FORWARD_FINALIZE)),
                "bwd_tile": alone(lambda: step.backward(None, _lib.STAGE_BACKWARD_TILE)),
                "bwd_fin": alone(lambda: step.backward(None, _lib.STAGE_BACKWARD_FINALIZE)),
            }
            print(f"loss {kind} {str(dtype):15s} tau {tau}: fused step median {ts[len(ts) // 2]:7.1f} us (min {ts[0]:.1f}) | "
                  + " ".join(f"{k} {v:6.1f}" for k, v in parts.items()) + f" | acc rows {float(step.stats[2]):.0f} loss {float(step.loss):.5f}",
                  flush=True)
            del gr, step
