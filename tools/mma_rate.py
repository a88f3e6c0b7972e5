"""tcgen05.mma issue / execution rate probe: cycles per 128xNx16 bf16 MMA under different kinds of contention."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
names = ["SS N=128", "SS N=256", "TS N=128 (A in TMEM)", "SS/TS N=128 alternating"]
modes = ["idle", "tcgen05.ld", "MUFU.EX2", "FFMA", "ld+MUFU+FFMA"]
sink = torch.zeros(640, device="cuda")
for mode, mname in enumerate(modes):
    out = torch.zeros(4 * 4, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _lib.check(lib.simclr_debug_mma_rate(out.data_ptr(), 64, 148, mode, sink.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "rate")
    torch.cuda.synchronize()
    o = out.cpu().view(4, 4)
    for i, n in enumerate(names):
        nmma = 8 * int(o[i, 2])
        print(f"16 warps {mname:14s} | {n:26s}: issue {int(o[i,0])/nmma:7.1f} cyc/MMA, complete {int(o[i,1])/nmma:7.1f} cyc/MMA")
