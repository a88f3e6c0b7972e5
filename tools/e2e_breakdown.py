import sys, time, torch
sys.path.insert(0, '.')
import pytorch_simclr_b200 as sb
sb.set_precision("bf16")
dev = torch.device("cuda", 0)
b, d = 4096, 128
g = torch.Generator().manual_seed(0)
h1 = torch.randn(b, d, generator=g).pin_memory(); h2 = torch.randn(b, d, generator=g).pin_memory()
def t(name, fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name:50s} {(time.perf_counter()-t0)/n*1e6:8.1f} us")
def copies():
    a = h1.to(dev, non_blocking=True); c = h2.to(dev, non_blocking=True); return a, c
t("2 H2D copies (async, no sync)", copies)
t("2 H2D copies + sync", lambda: (copies(), torch.cuda.synchronize()))
a, c = copies(); a.requires_grad_(True); c.requires_grad_(True)
for eager in (True, False):
    sb.set_eager_backward(eager)
    t(f"eager={eager}: contrastive_loss (fwd + acc item)", lambda: sb.contrastive_loss(a, c, temperature=0.5))
    def full():
        loss, acc = sb.contrastive_loss(a, c, temperature=0.5); loss.backward(); return loss.item()
    t(f"eager={eager}: loss + backward + item (device inputs)", full)
    def e2e():
        x, y = copies(); x.requires_grad_(True); y.requires_grad_(True)
        loss, acc = sb.contrastive_loss(x, y, temperature=0.5); loss.backward(); return loss.item()
    t(f"eager={eager}: e2e from pinned host", e2e)
from pytorch_simclr_b200 import functional as F
t("run_fused only (no sync)", lambda: F.run_fused(0, a, c, 0.5, True))
t("run_fused + stats.item", lambda: F.run_fused(0, a, c, 0.5, True)[1][2].item())
