"""Global-batch contrastive loss across the GPUs of one node (one process per GPU, NCCL over NVLink).

The reference is single-process; its only batch-scaling device is gradient accumulation, which does not
enlarge the set of negatives (utils/model_utils.py:113-123).  Here the 2B x 2B problem is sharded by ROWS:

  rank k holds its own images' two views  x1_k, x2_k  [B/R, d]   (so every positive pair is rank-local)
  forward : normalise locally -> all-gather the bf16 operands (2 collectives, one per view, into the
            view-padded global layout) -> each rank runs the tile kernel on its rows x all columns ->
            all-reduce of three scalars (sum w L, sum w, #correct)
  backward: all-gather of the per-row lse2 (M floats); the kernel uses the symmetric form
            W[r,c] = g_r P[r,c] + g_c P[c,r], so each rank produces the COMPLETE gradient of the global
            loss with respect to its own rows -- there is no column-partial reduce-scatter to do.

Every rank must hold the same number of images.  The returned loss is the global loss (identical on all
ranks) and each rank receives the COMPLETE gradient of that global loss with respect to its own rows.

Composition with DistributedDataParallel: DDP AVERAGES parameter gradients over the ranks.  Here every rank already
back-propagates the gradient of the whole global loss through its own rows, so the sum over ranks -- not the mean --
is the single-process global-batch gradient; under DDP the effective gradient would be 1/world of it.  Pass
``ddp_scale=True`` (multiplies the returned loss by the world size, so that DDP's mean is the exact gradient; the
value reported to the caller is then world x the global loss) or scale the loss yourself.
tests/distributed_check.py checks this against a single-process run.

Two transports:
  * ``PeerBatch`` (default on GPUs): the exchange is fused into our own kernels over peer memory.  The operand rows
    are stored by the prepare kernel straight into every rank's copy of the global operand matrix (symmetric
    memory, NVLink / NVSwitch), the forward finalize kernel stores lse2 and the loss statistics the same way, and
    two device-side flag barriers order producers and consumers.  No NCCL call on the data path, CUDA-graph capturable.
  * ``RowShardGather``: the same exchange with ``torch.distributed`` collectives (NCCL on GPUs, gloo on CPU tensors --
    which is how tests/test_distributed_cpu.py covers the layout logic without a GPU).  Used when per-row weights
    are given, and as the baseline the fused path is measured against.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

import ctypes

from . import _lib
from ._lib import check
from .functional import (LOSS_MODIFIED, LOSS_NTXENT, ContrastiveLossFunction, _Saved, _dtype_code, _validate, pad_dim,
                         pad_rows)

__all__ = ["RowShardGather", "PeerBatch", "global_contrastive_loss", "global_modified_contrastive_loss", "shard_rows"]


def shard_rows(b_global: int, world: int, rank: int):
    """(row_offset, b_local) of rank's contiguous image shard; the batch must divide evenly."""
    if b_global % world:
        raise ValueError(f"global batch {b_global} does not divide over {world} ranks")
    b_local = b_global // world
    return rank * b_local, b_local


class RowShardGather:
    """Collective plumbing between the local and the global view-padded layouts.

    Device-agnostic (works on CPU tensors with gloo, which is how tests/test_distributed_cpu.py covers it).
    """

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def _gather_views(self, local: torch.Tensor, b: int) -> torch.Tensor:
        """local: [P*2*pad(b), ...] view-padded (P operand planes: 1, or 2 in split precision)  ->
        [P*2*pad(b*world), ...] view-padded, rank-major inside a view."""
        bl_pad = pad_rows(b)
        planes = local.shape[0] // (2 * bl_pad)
        bg = b * self.world
        bg_pad = pad_rows(bg)
        out = local.new_zeros((planes * 2 * bg_pad,) + tuple(local.shape[1:]))
        for v in range(2 * planes):
            dist.all_gather_into_tensor(out[v * bg_pad: v * bg_pad + bg], local[v * bl_pad: v * bl_pad + b].contiguous(),
                                        group=self.group)
        return out

    # -- hooks used by functional.run_forward ------------------------------------------------
    def operand(self, operand_local: torch.Tensor, b: int):
        return self._gather_views(operand_local, b), b * self.world, self.rank * b

    def rowvec(self, vec_local: torch.Tensor, b: int) -> torch.Tensor:
        return self._gather_views(vec_local, b)

    def zrows(self, zrows_local: torch.Tensor, b: int) -> torch.Tensor:
        """The ranks' exact fp32 normalised rows [2*pad(b), Dpad] -> [2*pad(b*world), Dpad]: what the forward finalize
        kernel re-scores accuracy candidates from (include/simclr_b200.h, zrows_global)."""
        return self._gather_views(zrows_local, b)

    def reduce(self, stats: torch.Tensor, loss: torch.Tensor):
        """Sum the per-rank [sum w L, sum w, #correct]; the global loss is their ratio."""
        tot = stats[:3].clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        g_loss = tot[0] / tot[1]          # fresh 0-d tensor (not a view: callers divide it in place)
        g_stats = torch.cat((tot, g_loss.reshape(1)))
        return g_loss, g_stats

    def col_scale(self, w_local: torch.Tensor, g_stats: torch.Tensor, b: int) -> torch.Tensor:
        """w_c / sum(w) for every global column, view-padded."""
        bl_pad = pad_rows(b)
        padded = w_local.new_zeros(2 * bl_pad)
        padded[:b] = w_local[:b]
        padded[bl_pad:bl_pad + b] = w_local[b:]
        return self._gather_views(padded, b) / g_stats[1]


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerBatch:
    """Symmetric-memory state of the row-sharded global batch for one (b_local, d) on this rank.

    One symmetric allocation per rank holds two generations (double buffering: a fast rank may already push step
    k+1 while a slow one still reads step k) of  operand_all bf16 [2*Bgpad][Dpad] | colvec f32 [2][2*Bgpad] (the backward's column
    vectors a_c | lse2_c) | stats_all f32 [world][4]  plus the barrier flags u32 [world].  Everything else is ordinary device memory.
    """
    peer = True
    GENERATIONS = 2

    def __init__(self, b_local: int, d: int, group=None, device=None, use_multicast: bool = True,
                 overlap_local_first: bool = False):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 16:
            raise ValueError("PeerBatch supports up to 16 ranks of one node")
        self.b_local, self.d = int(b_local), int(d)
        self.b_global = self.b_local * self.world
        self.row_offset = self.b_local * self.rank
        self.bg_pad, self.dp = pad_rows(self.b_global), pad_dim(d)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        op_bytes = _align(2 * self.bg_pad * self.dp * 2)
        lse_bytes = _align(2 * 2 * self.bg_pad * 4)
        st_bytes = _align(self.world * 4 * 4)
        # this rank's exact fp32 normalised rows [2*Blpad][Dpad]: never pushed, read by the other ranks' finalize kernels
        # for the few accuracy candidates they have to re-score
        z_bytes = _align(2 * pad_rows(self.b_local) * self.dp * 4)
        self._gen_bytes = op_bytes + lse_bytes + st_bytes + z_bytes
        flag_off = self.GENERATIONS * self._gen_bytes
        total = flag_off + _align(16 * 4)
        self.buf = symm.empty(total, dtype=torch.uint8, device=self.device)
        self.buf.zero_()                       # padding rows of operand_all and the flags start (and stay) zero
        torch.cuda.synchronize(self.device)
        self.handle = symm.rendezvous(self.buf, self.group)
        dist.barrier(group=self.group)         # every rank has zeroed its copy before anybody pushes
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        arr = ctypes.c_void_p * self.world
        self._tables = []
        for g in range(self.GENERATIONS):
            base = g * self._gen_bytes
            self._tables.append({
                "operand": arr(*[p + base for p in ptrs]),
                "colvec": arr(*[p + base + op_bytes for p in ptrs]),
                "stats": arr(*[p + base + op_bytes + lse_bytes for p in ptrs]),
                "zrows": arr(*[p + base + op_bytes + lse_bytes + st_bytes for p in ptrs]),
            })
        self._flags = arr(*[p + flag_off for p in ptrs])
        # NVLS multicast mapping of the same allocation (0 when the fabric has no multicast support)
        mc = 0
        if use_multicast:
            try:
                mc = int(self.handle.multicast_ptr or 0)
            except Exception:       # handle without multicast support
                mc = 0
        self.multicast = bool(mc)
        self._mc = [None if not mc else mc + g * self._gen_bytes for g in range(self.GENERATIONS)]
        local = self.buf
        self._views = []
        for g in range(self.GENERATIONS):
            base = g * self._gen_bytes
            self._views.append({
                "operand": local[base: base + 2 * self.bg_pad * self.dp * 2].view(torch.bfloat16).view(2 * self.bg_pad, self.dp),
                "colvec": local[base + op_bytes: base + op_bytes + 2 * 2 * self.bg_pad * 4].view(torch.float32),
                "stats": local[base + op_bytes + lse_bytes: base + op_bytes + lse_bytes + self.world * 16].view(torch.float32),
            })
        # overlap_local_first: run the tiles of this rank's own columns before the barrier that waits for the peers'
        # operand rows (two tile-kernel launches).  Measured on 2/4/8 B200 at 2N = 65536: 1.444 / 0.7445 / 0.4132 ms
        # against 1.422 / 0.7396 / 0.4111 ms without -- the NVLS push is not what limits the step, the second launch's
        # fixed cost is slightly more than the hidden transfer -- hence off by default.
        self.overlap = bool(overlap_local_first)
        self.epoch = torch.zeros(4, dtype=torch.int32, device=self.device)      # [0] epoch counter, [1] warp ticket of the fused step's prepare kernel
        self.generation = 0                    # number of forwards issued so far
        self.lib = _lib.load()

    # ------------------------------------------------------------------------------------------------------
    def forward(self, loss_kind, x1, x2, temperature, normalize, operand, rowvec, stats_local, stats_global, loss,
                ws, ws_bytes, stream, bwd_ws=None):
        """prepare(+push) -> barrier -> forward(+push) -> barrier(+global statistics); enqueue only."""
        lib, gen = self.lib, self.generation % self.GENERATIONS
        tab, view = self._tables[gen], self._views[gen]
        code = _dtype_code(x1)
        check(lib.simclr_prepare_peer(loss_kind, x1.data_ptr(), x2.data_ptr(), self.b_local, self.d, code,
                                      int(bool(normalize)), float(temperature), _lib.PRECISION_BF16, operand.data_ptr(),
                                      rowvec[0].data_ptr(), rowvec[1].data_ptr(), ws.data_ptr(), self.world, self.rank,
                                      tab["operand"], self._mc[gen], tab["zrows"][self.rank], stream),
              "simclr_prepare_peer")
        if not self.overlap:
            check(lib.simclr_peer_barrier(self.world, self.rank, self._flags, self.epoch.data_ptr(), None, None, None,
                                          stream), "simclr_peer_barrier")
        # (overlap: the barrier between the operand push and its consumers is issued inside simclr_forward_peer, after
        # the tiles of this rank's own columns, which overlap the NVLink transfer of everybody else's rows)
        check(lib.simclr_forward_peer(loss_kind, operand.data_ptr(), view["operand"].data_ptr(), self.b_local,
                                      self.b_global, self.row_offset, self.d, float(temperature), int(bool(normalize)),
                                      _lib.PRECISION_BF16, rowvec[1].data_ptr(), None, rowvec[2].data_ptr(),
                                      rowvec[3].data_ptr(),
                                      stats_local.data_ptr(), None, ws.data_ptr(), ws_bytes,
                                      None if bwd_ws is None else bwd_ws.data_ptr(), 0 if bwd_ws is None else bwd_ws.numel(),
                                      self.world, self.rank, tab["colvec"], tab["stats"],
                                      self._flags if self.overlap else None,
                                      self.epoch.data_ptr() if self.overlap else None, None, None, code, None, None,
                                      tab["zrows"], stream), "simclr_forward_peer")
        check(lib.simclr_peer_barrier(self.world, self.rank, self._flags, self.epoch.data_ptr(), view["stats"].data_ptr(),
                                      stats_global.data_ptr(), loss.data_ptr(), stream), "simclr_peer_barrier")
        self.generation += 1
        return view["operand"], view["colvec"], self.generation


    def fused_step(self, loss_kind, x1, x2, temperature, normalize, operand, rowvec, stats_local, stats_global, loss,
                   grad1, grad2, fwd_ws, fwd_ws_bytes, bwd_ws, bwd_ws_bytes, stream, grad_out=None):
        """The whole row-sharded step of this rank in five launches (simclr_forward_backward_peer): the two cross-GPU
        barriers run inside the tile kernels.  Enqueue only; uses the next buffer generation like forward()."""
        gen = self.generation % self.GENERATIONS
        tab = self._tables[gen]
        check(self.lib.simclr_forward_backward_peer(
            loss_kind, x1.data_ptr(), x2.data_ptr(), self.b_local, self.d, _dtype_code(x1), int(bool(normalize)),
            float(temperature), None if grad_out is None else grad_out.data_ptr(), operand.data_ptr(), rowvec.data_ptr(),
            stats_local.data_ptr(), stats_global.data_ptr(), loss.data_ptr(), grad1.data_ptr(), grad2.data_ptr(),
            fwd_ws.data_ptr(), fwd_ws_bytes, bwd_ws.data_ptr(), bwd_ws_bytes, self.world, self.rank, tab["operand"],
            self._mc[gen], tab["colvec"], tab["stats"], self._flags, self.epoch.data_ptr(), tab["zrows"], 0, stream),
            "simclr_forward_backward_peer")
        self.generation += 1
        return self.generation


def run_forward_peer(loss_kind: int, x1: torch.Tensor, x2: torch.Tensor, temperature: float, normalize: bool,
                     peer: PeerBatch, prime_backward: bool = True):
    """Peer-memory counterpart of functional.run_forward (same return tuple)."""
    lib = _lib.load()
    b, d = _validate(x1, x2)
    if (b, d) != (peer.b_local, peer.d):
        raise ValueError(f"PeerBatch was built for batches of {peer.b_local} x {peer.d}, got {b} x {d}")
    x1, x2 = x1.contiguous(), x2.contiguous()
    dev = x1.device
    bp, dp = pad_rows(b), pad_dim(d)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        operand = torch.empty((2 * bp, dp), dtype=torch.bfloat16, device=dev)
        rowvec = torch.empty((4, 2 * bp), dtype=torch.float32, device=dev)
        stats_local = torch.empty(4, dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws_bytes = lib.simclr_forward_workspace_bytes(loss_kind, b, peer.b_global, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        bwd_ws = None
        if prime_backward:
            bwd_ws = torch.empty(lib.simclr_backward_workspace_bytes(loss_kind, b, peer.b_global, d), dtype=torch.uint8,
                                 device=dev)
        operand_cols, colvec, generation = peer.forward(loss_kind, x1, x2, temperature, normalize, operand, rowvec,
                                                        stats_local, stats, loss, ws, ws_bytes, stream, bwd_ws)
    saved = _Saved()
    saved.operand_rows, saved.operand_cols = operand, operand_cols
    saved.inv_norm, saved.pos_dot = rowvec[0], rowvec[1]
    saved.lse2_cols, saved.col_scale = colvec[2 * peer.bg_pad:], None      # plane 1 of the column vectors is lse2
    saved.bwd_ws = bwd_ws
    saved.primed_colvec = colvec.data_ptr() if bwd_ws is not None else None
    saved.b_local, saved.b_global, saved.row_offset, saved.d = b, peer.b_global, peer.row_offset, d
    saved.loss, saved.temperature, saved.normalize, saved.dtype_code = loss_kind, float(temperature), bool(normalize), _dtype_code(x1)
    saved.peer, saved.generation = peer, generation
    saved.precision = _lib.PRECISION_BF16
    return loss, stats, rowvec, saved


_PEER_CACHE = {}


def _peer_for(x1: torch.Tensor, group) -> PeerBatch:
    key = (id(group), x1.device, x1.shape[0], x1.shape[1])
    if key not in _PEER_CACHE:
        _PEER_CACHE[key] = PeerBatch(x1.shape[0], x1.shape[1], group, x1.device)
    return _PEER_CACHE[key]


def _global_loss(kind, x1, x2, temperature, normalize, weight, group, transport="auto", ddp_scale=False):
    from .functional import get_precision
    if transport not in ("auto", "peer", "nccl"):
        raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
    # The peer-memory transport carries bf16 operands and the unweighted loss only: anything else must not be dropped
    # silently.  "auto" routes such calls through the collectives; an explicit "peer" is an error.
    peer_ok = weight is None and get_precision() != "fp32" and x1.is_cuda
    if transport == "peer" and not peer_ok:
        if weight is not None:
            raise ValueError("transport='peer' does not support per-row weights: use transport='nccl' (or 'auto')")
        if not x1.is_cuda:
            raise ValueError("transport='peer' needs CUDA tensors")
        raise ValueError("transport='peer' computes with bf16 tensor-core operands; precision 'fp32' was requested "
                         "(set_precision): use transport='nccl' (or 'auto')")
    use_peer = transport == "peer" or (transport == "auto" and peer_ok)
    if use_peer:
        gather = _peer_for(x1, group)
        weight = None
    else:
        gather = RowShardGather(group)
    loss, stats = ContrastiveLossFunction.apply(x1, x2, kind, float(temperature), bool(normalize), weight, gather)
    correct = stats[2].item()
    if ddp_scale:
        loss = loss * gather.world
    return loss, 100.0 * correct / (2 * x1.shape[0] * gather.world)


def global_contrastive_loss(x_batch1, x_batch2, temperature=1.0, normalize=True, weight: Optional[torch.Tensor] = None,
                            group=None, transport: str = "auto", ddp_scale: bool = False):
    """NT-Xent over the union of all ranks' batches.  Same signature and return convention as
    ``contrastive_loss`` (reference objective.py:6-10,55); ``weight`` is this rank's [2*B_local] slice.
    ``transport``: "peer" (fused NVLink stores + device barriers; unweighted, bf16 operands -- anything else raises),
    "nccl" (collectives) or "auto".  ``ddp_scale``: see the module docstring (gradient averaging under DDP)."""
    return _global_loss(LOSS_NTXENT, x_batch1, x_batch2, temperature, normalize, weight, group, transport, ddp_scale)


def global_modified_contrastive_loss(x_batch1, x_batch2, group=None, transport: str = "auto", ddp_scale: bool = False,
                                     **kwargs):
    """Probabilistic loss over the union of all ranks' batches (reference objective.py:58-98)."""
    return _global_loss(LOSS_MODIFIED, x_batch1, x_batch2, kwargs.get("temperature", 1.0), True, None, group, transport,
                        ddp_scale)
