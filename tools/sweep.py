"""Loss sweep of BASELINE.json configs[4]: NT-Xent fwd+bwd over 2N = 1024 ... 131072, d in {128, 256}, tau in {0.1, 0.5}
(bf16 mode; the fp32-grade mode where it applies), one GPU, CUDA-graph replay with L2 flush, CUDA events.
Output is committed as profiles/r01_sweep.log."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200.functional import LOSS_MODIFIED, LOSS_NTXENT  # noqa: E402
from pytorch_simclr_b200.runner import ContrastiveStep  # noqa: E402

peak = 1665.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.isfile(p):
    peak = float(json.load(open(p))["bf16_tflops"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()


def time_step(kind, b, d, tau, precision, dtype=torch.float32):
    step = ContrastiveStep(kind, b, d, tau, True, dtype, "cuda", precision)
    g = torch.Generator().manual_seed(b + d)
    step.x1.copy_(torch.randn(b, d, generator=g))
    step.x2.copy_(torch.randn(b, d, generator=g))
    with torch.cuda.stream(side):
        for _ in range(2):
            step.step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step.step()
    n = 20 if b <= 8192 else (8 if b <= 32768 else 4)
    for _ in range(2):
        flush.zero_()
        graph.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, z in ev:
        flush.zero_()
        a.record()
        graph.replay()
        z.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(z) for a, z in ev) / n
    return ms, float(step.loss)


print(f"{'loss':9s} {'2N':>7s} {'d':>4s} {'tau':>4s} {'mode':5s} | {'ms/step':>9s} {'Mviews/s':>9s} {'TFLOP/s (6M^2d | 3M^2d)':>24s} {'of measured bf16 peak':>22s} | loss")
for kind, name, flop_c in ((LOSS_NTXENT, "ntxent", 6.0), (LOSS_MODIFIED, "modified", 3.0)):
    for d in (128, 256):
        for tau in (0.5, 0.1):
            for m in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072):
                if name == "modified" and (d == 256 or m in (2048, 4096, 16384, 32768)):
                    continue
                for mode in (("bf16", "fp32") if (d <= 128 and m in (1024, 8192, 65536) and tau == 0.5) else ("bf16",)):
                    ms, loss = time_step(kind, m // 2, d, tau, mode)
                    tf = flop_c * m * m * d / (ms * 1e-3) / 1e12
                    print(f"{name:9s} {m:7d} {d:4d} {tau:4.1f} {mode:5s} | {ms:9.4f} {m / ms / 1e3:9.2f} {tf:24.1f} {tf / peak:22.3f} | {loss:.5f}",
                          flush=True)
