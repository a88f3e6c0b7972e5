"""Kernel-level %globaltimer timeline of the row-sharded global step (PeerStep) on every rank, launched with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/peer_timeline.py [--b 32768] [--d 128]

Prints per rank the start / end of prepare, forward tile, forward finalize, backward tile, backward finalize relative
to the rank's own prepare start; the gaps prepare -> forward tile and forward finalize -> backward tile are the two
device-side cross-GPU barriers (waiting for the slowest rank and for the NVLink pushes included)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the stamps only exist in the tracing build: make it the library the whole package uses in this process
os.environ.setdefault("SIMCLR_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     "pytorch-simclr_b200", "lib", "libsimclr_b200_trace.so"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pytorch_simclr_b200 import _lib  # noqa: E402
from pytorch_simclr_b200.distributed import shard_rows  # noqa: E402
from pytorch_simclr_b200.runner import PeerStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=32768)
ap.add_argument("--d", type=int, default=128)
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
off, bl = shard_rows(args.b, world, rank)
step = PeerStep(0, bl, args.d, 0.5, None, True, torch.float32, dev)
gen = torch.Generator().manual_seed(1000 + rank)
step.x1.copy_(torch.randn(bl, args.d, generator=gen))
step.x2.copy_(torch.randn(bl, args.d, generator=gen))
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib.simclr_debug_set_kernel_trace(buf.data_ptr())
side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    for _ in range(3):
        step.step()
torch.cuda.synchronize()
dist.barrier()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    step.step()
torch.cuda.synchronize()
dist.barrier()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = {0: "prepare", 1: "fwd_tile", 4: "fwd_fin", 3: "bwd_tile", 5: "bwd_fin"}
lines = []
for rep in range(4):
    v = torch.zeros(16, dtype=torch.int64)
    v[0::2] = torch.iinfo(torch.int64).max
    buf.copy_(v)
    flush.zero_()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    graph.replay()
    ev1.record()
    torch.cuda.synchronize()
    t = buf.cpu().view(8, 2)
    t0 = int(t[0, 0])
    seg = " | ".join(f"{names[k]} {(int(t[k, 0]) - t0) / 1e3:7.1f}..{(int(t[k, 1]) - t0) / 1e3:7.1f}" for k in (0, 1, 4, 3, 5))
    lines.append(f"[rank {rank} rep {rep}] events {ev0.elapsed_time(ev1) * 1e3:7.1f} us | {seg}")
lib.simclr_debug_set_kernel_trace(None)
for r in range(world):
    dist.barrier()
    if r == rank:
        print("\n".join(lines[1:]), flush=True)
dist.destroy_process_group()
