"""A/B timing of kernel variants: fused step (CUDA graph, L2 flushed), forward stage and backward stage, mean over
`--reps` replays, for the library selected by SIMCLR_B200_LIB (see tools/variants.sh)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=4096)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--loss", type=int, default=0)
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--data", default="iid", choices=["iid", "correlated"])
args = ap.parse_args()
step = ContrastiveStep(args.loss, args.b, args.d, 0.5)
g = torch.Generator().manual_seed(0)
if args.data == "iid":
    step.x1.copy_(torch.randn(args.b, args.d, generator=g))
    step.x2.copy_(torch.randn(args.b, args.d, generator=g))
else:
    base = torch.randn(args.b, args.d, generator=g)
    step.x1.copy_(base + 0.5 * torch.randn(args.b, args.d, generator=g))
    step.x2.copy_(base + 0.5 * torch.randn(args.b, args.d, generator=g))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        step.step()
        step.step_staged()
torch.cuda.synchronize()
graphs = {}
for name, fn in (("step", step.step), ("fwd", step.forward), ("bwd", step.backward)):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        fn()
    graphs[name] = gr
# eight fused steps back to back in one graph (programmatic launch across step boundaries)
g8 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g8, stream=side):
    for _ in range(8):
        step.step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(gr, reps, div=1):
    for _ in range(5):
        flush.zero_()
        gr.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()
        a.record()
        gr.replay()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ms) / len(ms) * 1e3 / div, ms[len(ms) // 10] * 1e3 / div


out = []
for name in ("step", "fwd", "bwd"):
    mean, p10 = t(graphs[name], args.reps)
    out.append(f"{name} {mean:6.1f} (p10 {p10:6.1f})")
mean, p10 = t(g8, max(args.reps // 4, 10), 8)
out.append(f"step x8 in one graph {mean:6.1f} (p10 {p10:6.1f})")
print(" | ".join(out) + f" us | loss {float(step.loss):.6f}")
