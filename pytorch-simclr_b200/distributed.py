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
ranks); its gradient w.r.t. the local inputs is exact, so DDP-style parameter all-reduce composes as usual.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .functional import LOSS_MODIFIED, LOSS_NTXENT, ContrastiveLossFunction, pad_rows

__all__ = ["RowShardGather", "global_contrastive_loss", "global_modified_contrastive_loss", "shard_rows"]


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
        """local: [2*pad(b), ...] view-padded  ->  [2*pad(b*world), ...] view-padded, rank-major inside a view."""
        bl_pad = local.shape[0] // 2
        bg = b * self.world
        bg_pad = pad_rows(bg)
        out = local.new_zeros((2 * bg_pad,) + tuple(local.shape[1:]))
        for v in (0, 1):
            dist.all_gather_into_tensor(out[v * bg_pad: v * bg_pad + bg], local[v * bl_pad: v * bl_pad + b].contiguous(),
                                        group=self.group)
        return out

    # -- hooks used by functional.run_forward ------------------------------------------------
    def operand(self, operand_local: torch.Tensor, b: int):
        return self._gather_views(operand_local, b), b * self.world, self.rank * b

    def rowvec(self, vec_local: torch.Tensor, b: int) -> torch.Tensor:
        return self._gather_views(vec_local, b)

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


def _global_loss(kind, x1, x2, temperature, normalize, weight, group):
    gather = RowShardGather(group)
    loss, stats = ContrastiveLossFunction.apply(x1, x2, kind, float(temperature), bool(normalize), weight, gather)
    correct = stats[2].item()
    return loss, 100.0 * correct / (2 * x1.shape[0] * gather.world)


def global_contrastive_loss(x_batch1, x_batch2, temperature=1.0, normalize=True, weight: Optional[torch.Tensor] = None,
                            group=None):
    """NT-Xent over the union of all ranks' batches.  Same signature and return convention as
    ``contrastive_loss`` (reference objective.py:6-10,55); ``weight`` is this rank's [2*B_local] slice."""
    return _global_loss(LOSS_NTXENT, x_batch1, x_batch2, temperature, normalize, weight, group)


def global_modified_contrastive_loss(x_batch1, x_batch2, group=None, **kwargs):
    """Probabilistic loss over the union of all ranks' batches (reference objective.py:58-98)."""
    return _global_loss(LOSS_MODIFIED, x_batch1, x_batch2, kwargs.get("temperature", 1.0), True, None, group)
