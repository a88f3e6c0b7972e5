"""Softmax-side throughput probe: cycles per 32-column chunk for variants of the per-element arithmetic, with 4 / 8 /
16 active warps per SM.  A 128x128 tile costs each SM sub-partition 16 chunk-times (4 warps x 4 chunks...), i.e.
cycles/tile = 16 * (cycles per chunk with all 16 warps active) / 4 warps in flight."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
names = ["fwd shipped (1/4 poly4)", "bwd shipped (1/4 poly3)", "fwd all-MUFU + max", "fwd all-MUFU no max", "fwd 1/2 poly4",
         "MUFU + FADD only", "FFMA only (4/elem)", "bwd all-MUFU", "tcgen05.ld only", "fwd 1/4 poly3"]
sink = torch.zeros(640, device="cuda")
for nwarps in (4, 8, 16):
    out = torch.zeros(10 * 32, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _lib.check(lib.simclr_debug_chunk_rate(out.data_ptr(), 2000, 148, nwarps, 2.885, sink.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "chunk_rate")
    torch.cuda.synchronize()
    o = out.cpu().view(10, 32)[:, :nwarps].double()
    for i, n in enumerate(names):
        per_half = o[i].mean().item()
        # a 128x128 tile = 8 warp-level half tiles (4 lane quarters x 2 column halves) = 2 per SM sub-partition; with
        # w = nwarps/4 warps per sub-partition in flight it finishes w half tiles per `per_half` cycles
        per_tile = 2 * per_half / (nwarps / 4)
        print(f"{nwarps:2d} warps | {n:26s}: {per_half:7.1f} cyc / 64-col half tile / warp  -> {per_tile:7.0f} cyc per 128x128 tile (softmax side only)")
