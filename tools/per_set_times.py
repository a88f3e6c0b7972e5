"""Per-input-set step times of bench.py's headline protocol: does a step's duration depend on how many rows the accuracy
count had to re-score exactly (forward workspace header: word 1 = listed rows, word 2 = confirmed hits)?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

b, d = 4096, 128
step = ContrastiveStep(0, b, d, 0.5)
dgen = torch.Generator(device="cuda").manual_seed(1)
sets = [(torch.randn(b, d, generator=dgen, device="cuda"), torch.randn(b, d, generator=dgen, device="cuda"),
         torch.empty(b, d, device="cuda"), torch.empty(b, d, device="cuda")) for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16)]
side = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
hdr = step.fwd_ws[:16].view(torch.int32)
for i, (x1, x2, g1, g2) in enumerate(sets):
    with torch.cuda.stream(side):
        step.step(None, x1, x2, g1, g2)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step.step(None, x1, x2, g1, g2)
    ts = []
    for _ in range(5):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z) * 1e3)
    h = hdr.cpu().tolist()
    print(f"set {i:2d}: {min(ts):6.1f} us (median {sorted(ts)[2]:6.1f})  listed rows {h[1]:3d}  confirmed {h[2]:3d}  correct rows {float(step.stats[2]):.0f}")
