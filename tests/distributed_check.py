"""Row-sharded global-batch check, launched with torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/distributed_check.py [--batch 2048] [--dim 128]

Every rank compares the global loss / accuracy and its local gradients with the single-process fp64 oracle
evaluated on the gathered batch (test infrastructure: imports oracle/)."""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import contrastive_oracle as oracle  # noqa: E402
from pytorch_simclr_b200.distributed import (global_contrastive_loss, global_modified_contrastive_loss,  # noqa: E402
                                             shard_rows)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", dest="b", type=int, default=2048)
ap.add_argument("--dim", dest="d", type=int, default=128)
ap.add_argument("--tau", type=float, default=0.5)
args = ap.parse_args()

os.environ.setdefault("SIMCLR_B200_PEER_TIMEOUT_S", "120")      # a test's ranks run in lock step: do not wait ten minutes for a dead one
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, fn, ref_fn, use_weight, transport in (
        ("ntxent/peer", global_contrastive_loss, oracle.ntxent_closed_form, False, "peer"),
        ("ntxent/peer again", global_contrastive_loss, oracle.ntxent_closed_form, False, "peer"),
        ("ntxent/peer overlap", global_contrastive_loss, oracle.ntxent_closed_form, False, "peer-overlap"),
        ("ntxent/nccl", global_contrastive_loss, oracle.ntxent_closed_form, False, "nccl"),
        ("ntxent+weight/nccl", global_contrastive_loss, oracle.ntxent_closed_form, True, "nccl"),
        ("modified/peer", global_modified_contrastive_loss, oracle.modified_closed_form, False, "peer"),
        ("modified/nccl", global_modified_contrastive_loss, oracle.modified_closed_form, False, "nccl")):
    z1, z2 = oracle.make_embeddings(args.b, args.d, seed=17, kind="correlated", noise=1.0)   # same on every rank
    w = None
    if use_weight:
        w = torch.rand(2 * args.b, generator=torch.Generator().manual_seed(3)) + 0.25
    off, bl = shard_rows(args.b, world, rank)
    a = z1[off:off + bl].cuda().requires_grad_(True)
    c = z2[off:off + bl].cuda().requires_grad_(True)
    kw = dict(temperature=args.tau, transport=transport)
    if transport == "peer-overlap":
        # the overlapped schedule (local-column tiles, barrier, remote-column tiles) through its own PeerBatch
        from pytorch_simclr_b200 import distributed as D
        from pytorch_simclr_b200.functional import ContrastiveLossFunction, LOSS_NTXENT
        if "ov" not in globals():
            ov = D.PeerBatch(args.b // world, args.d, None, torch.device("cuda", local), overlap_local_first=True)

        def fn(x1, x2, temperature, transport):      # noqa: F811
            loss, stats = ContrastiveLossFunction.apply(x1, x2, LOSS_NTXENT, float(temperature), True, None, ov)
            return loss, 100.0 * stats[2].item() / (2 * x1.shape[0] * world)
    if use_weight:
        kw["weight"] = torch.cat((w[off:off + bl], w[args.b + off:args.b + off + bl])).cuda()
    loss, acc = fn(a, c, **kw)
    loss.backward()
    torch.cuda.synchronize()
    ref = ref_fn(z1, z2, temperature=args.tau, **({"weight": w.numpy()} if use_weight else {}))
    gmax = max(np.abs(ref.grad1).max(), np.abs(ref.grad2).max())
    e1 = np.abs(a.grad.cpu().numpy() - ref.grad1[off:off + bl]).max() / gmax
    e2 = np.abs(c.grad.cpu().numpy() - ref.grad2[off:off + bl]).max() / gmax
    lrel = abs(float(loss.detach()) - ref.loss) / abs(ref.loss)
    # the accuracy count is exact on every transport: near-ties inside bf16 resolution are re-scored in exact fp32 from
    # the owning rank's rows (symmetric memory / gathered copy)
    good = lrel < 2e-3 and e1 < 1e-2 and e2 < 1e-2 and round(acc * 2 * args.b / 100.0) == ref.correct
    ok = ok and good
    print(f"[rank {rank}/{world}] {name}: loss {float(loss.detach()):.6f} (oracle {ref.loss:.6f}, rel {lrel:.1e}) "
          f"acc {acc:.3f} (oracle {ref.acc:.3f}) grad err {e1:.1e} {e2:.1e} -> {'OK' if good else 'FAIL'}", flush=True)
# ---- the fused row-sharded step (PeerStep.step: five launches, cross-GPU barriers inside the tile kernels): four steps
# over two different global batches, enqueued back to back without any host synchronisation in between ----
from pytorch_simclr_b200.runner import PeerStep  # noqa: E402
from pytorch_simclr_b200.functional import LOSS_MODIFIED, LOSS_NTXENT  # noqa: E402

for kind, ref_fn, label in ((LOSS_NTXENT, oracle.ntxent_closed_form, "ntxent"), (LOSS_MODIFIED, oracle.modified_closed_form, "modified")):
    off, bl = shard_rows(args.b, world, rank)
    ps = PeerStep(kind, bl, args.d, args.tau, None, True, torch.float32, torch.device("cuda", local))
    batches = [oracle.make_embeddings(args.b, args.d, seed=40 + s, kind="correlated" if s else "iid", noise=1.0) for s in range(2)]
    go = torch.tensor([0.25], device="cuda")
    outs = []
    for i in range(4):
        z1, z2 = batches[i % 2]
        x1, x2 = z1[off:off + bl].cuda(), z2[off:off + bl].cuda()
        g1, g2 = torch.empty_like(x1), torch.empty_like(x2)
        ps.step(go, x1, x2, g1, g2)
        outs.append((x1, x2, g1, g2, ps.stats.clone(), ps.loss.clone()))
    torch.cuda.synchronize()
    for i, (x1, x2, g1, g2, st, ls) in enumerate(outs):
        z1, z2 = batches[i % 2]
        ref = ref_fn(z1, z2, temperature=args.tau, grad_output=0.25)
        gmax = max(np.abs(ref.grad1).max(), np.abs(ref.grad2).max())
        e1 = np.abs(g1.cpu().numpy() - ref.grad1[off:off + bl]).max() / gmax
        e2 = np.abs(g2.cpu().numpy() - ref.grad2[off:off + bl]).max() / gmax
        lrel = abs(float(ls) - ref.loss) / abs(ref.loss)
        acc = 100.0 * float(st[2]) / (2 * args.b)
        good = lrel < 2e-3 and e1 < 1e-2 and e2 < 1e-2 and int(round(float(st[2]))) == ref.correct
        ok = ok and good
        print(f"[rank {rank}/{world}] {label}/fused step {i}: loss {float(ls):.6f} (oracle {ref.loss:.6f}, rel {lrel:.1e}) "
              f"acc {acc:.3f} (oracle {ref.acc:.3f}) grad err {e1:.1e} {e2:.1e} -> {'OK' if good else 'FAIL'}", flush=True)
    del ps
# ---- DistributedDataParallel training step with the global loss: a small encoder wrapped in DDP, every
# rank feeds its shard, loss = global_contrastive_loss(..., ddp_scale=True); after backward() the (DDP-averaged) parameter
# gradients must equal the gradients of the single-process loss over the WHOLE batch (fp64 reference: the same encoder
# applied to the gathered batch, oracle loss gradients chained through torch autograd) ----
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

torch.manual_seed(0)                                              # identical initial weights on every rank
enc = torch.nn.Sequential(torch.nn.Linear(64, 96), torch.nn.ReLU(), torch.nn.Linear(96, args.d, bias=False)).cuda()
ddp = DDP(enc, device_ids=[local])
gen = torch.Generator().manual_seed(5)
base = torch.randn(args.b, 64, generator=gen)
v1 = base + 0.3 * torch.randn(args.b, 64, generator=gen)
v2 = base + 0.3 * torch.randn(args.b, 64, generator=gen)
off, bl = shard_rows(args.b, world, rank)
z1 = ddp(v1[off:off + bl].cuda())
z2 = ddp(v2[off:off + bl].cuda())
loss, acc = global_contrastive_loss(z1, z2, temperature=args.tau, ddp_scale=True)
loss.backward()
torch.cuda.synchronize()
ref_enc = torch.nn.Sequential(torch.nn.Linear(64, 96), torch.nn.ReLU(), torch.nn.Linear(96, args.d, bias=False)).double()
ref_enc.load_state_dict({k: v.detach().cpu().double() for k, v in enc.state_dict().items()})
r1, r2 = ref_enc(v1.double()), ref_enc(v2.double())
ref = oracle.ntxent_closed_form(r1.detach(), r2.detach(), temperature=args.tau)
torch.autograd.backward([r1, r2], [torch.from_numpy(ref.grad1), torch.from_numpy(ref.grad2)])
worst = 0.0
for (name, p_), q in zip(enc.named_parameters(), ref_enc.parameters()):
    err = float((p_.grad.detach().cpu().double() - q.grad).abs().max() / q.grad.abs().max())
    worst = max(worst, err)
good = worst < 1e-2 and abs(float(loss.detach()) / world - ref.loss) / ref.loss < 2e-3 and round(acc * 2 * args.b / 100.0) == ref.correct
ok = ok and good
print(f"[rank {rank}/{world}] DDP step with ddp_scale: loss/world {float(loss.detach()) / world:.6f} (oracle {ref.loss:.6f}) "
      f"acc {acc:.3f} (oracle {ref.acc:.3f}) worst parameter-gradient error {worst:.1e} -> {'OK' if good else 'FAIL'}", flush=True)

flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
