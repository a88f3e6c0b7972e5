"""Issue rate of single opcodes on one SM sub-partition: cycles per warp instruction with 1, 2, 4 warps per
sub-partition (4, 8, 16 warps per SM), 16 independent chains per warp."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
names = ["FFMA reg,reg,reg", "FFMA reg,imm,reg", "FADD reg,reg", "FADD reg,imm", "FMUL reg,reg", "MUFU.EX2", "FMNMX reg,reg",
         "IMAD x*2^23+y", "SHL+IADD", "3 FFMA : 1 MUFU", "F2FP bf16x2", "FMNMX3", "LDS.128", "FSETP+FSEL", "HFMA2.BF16", "PRMT",
         "FADD2", "FFMA2", "FADD2 : FMNMX 1:1", "3 FFMA2 : 1 MUFU", "FADD2 : FADD 1:1", "EX2.F16x2", "EX2.BF16x2", "F2FP f16x2"]
sink = torch.zeros(640, device="cuda")
iters = 2000
for nwarps in (4, 8, 16):
    out = torch.zeros(32, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _lib.check(lib.simclr_debug_pipe_rate(out.data_ptr(), iters, 148, nwarps, sink.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "pipe_rate")
    torch.cuda.synchronize()
    o = out.cpu().double()
    for i, n in enumerate(names):
        ninstr = iters * 4 * 16 * (nwarps / 4)          # warp instructions issued on one sub-partition
        if i in (8, 13):
            ninstr *= 2
        if 16 <= i <= 20:
            ninstr /= 2                                 # eight (packed or scalar) instructions per 16-chain round
        print(f"{nwarps // 4} warps/SMSP | {n:18s}: {o[i].item() / ninstr:6.2f} cycles per warp instruction per sub-partition")
