for v in b200 v_nopp v_poly8 v_nopp_poly8 b200; do
  echo "== $v"
  SIMCLR_B200_LIB=$PWD/pytorch-simclr_b200/lib/libsimclr_$v.so python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from pytorch_simclr_b200.functional import LOSS_NTXENT
from pytorch_simclr_b200.runner import ContrastiveStep
step = ContrastiveStep(LOSS_NTXENT, 4096, 128, 0.5)
g = torch.Generator().manual_seed(0)
step.x1.copy_(torch.randn(4096, 128, generator=g)); step.x2.copy_(torch.randn(4096, 128, generator=g))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3): step.step()
torch.cuda.synchronize()
gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
with torch.cuda.graph(gf, stream=side): step.forward()
with torch.cuda.graph(gb, stream=side): step.backward()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(gr):
    for _ in range(5): flush.zero_(); gr.replay()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_(); a.record(); gr.replay(); b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return sum(ms) / len(ms) * 1e3
print("fwd stage %.1f us   bwd stage %.1f us   loss %.6f" % (t(gf), t(gb), float(step.loss)))
PY
done
