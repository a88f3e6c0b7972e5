"""Drop-in mirror of the reference's ``objective.py`` (same names, signatures and return conventions).

    contrastive_loss(x_batch1, x_batch2, temperature=1.0, normalize=True, weight=None) -> (loss, acc)
        reference objective.py:6-55
    modified_contrastive_loss(x_batch1, x_batch2, **kwargs) -> (loss, top1_acc)
        reference objective.py:58-98  (only ``temperature`` is read from kwargs, :68)

``loss`` is a 0-d float32 tensor on the inputs' device, connected to autograd; ``acc`` is a python float in
[0, 100] (this forces the same device->host sync the reference performs at objective.py:52 / :96).
The arithmetic runs in the sm_100a extension; CPU tensors raise ``ValueError``.
"""
from __future__ import annotations

import os

import torch

from . import _hoststats
from .functional import LOSS_MODIFIED, LOSS_NTXENT, ContrastiveLossFunction

__all__ = ["contrastive_loss", "modified_contrastive_loss", "set_lazy_accuracy", "get_lazy_accuracy"]

# The reference returns the accuracy as a python float (objective.py:52-53 / :96-97), i.e. every call ends with a
# device -> host read, and its training loop adds a second one (loss.item(), utils/model_utils.py:117).  With
# set_lazy_accuracy(True) the accuracy comes back as a 0-d device tensor instead and the call enqueues without any host
# synchronisation (what a CUDA-graph-captured or fully asynchronous training step needs); float(acc) gives the number.
_LAZY_ACCURACY = False


def set_lazy_accuracy(flag: bool) -> None:
    global _LAZY_ACCURACY
    _LAZY_ACCURACY = bool(flag)


def get_lazy_accuracy() -> bool:
    return _LAZY_ACCURACY


def _as_supported(x: torch.Tensor) -> torch.Tensor:
    # fp16 / fp64 inputs are computed in fp32 (autograd casts the gradients back)
    return x if x.dtype in (torch.float32, torch.bfloat16) else x.float()


# SIMCLR_B200_HOST_STATS=0 falls back to tensor.item() for the accuracy read-back (measurement / debugging)
_HOST_STATS = os.environ.get("SIMCLR_B200_HOST_STATS", "1") != "0"


def _loss_and_accuracy(x1, x2, kind, temperature, normalize, weight):
    """(loss tensor, accuracy float).  The accuracy forces a device->host read (reference objective.py:52 / :96); for the
    unweighted losses the finalize kernel writes the statistics into pinned host memory and this thread polls it
    (_hoststats.py) instead of paying a stream synchronisation."""
    if _LAZY_ACCURACY and x1.is_cuda:
        loss, stats = ContrastiveLossFunction.apply(x1, x2, kind, temperature, normalize, weight, None)
        return loss, stats[2] * (100.0 / (2 * x1.shape[0]))
    if weight is None and x1.is_cuda and _HOST_STATS:
        ring = _hoststats.ring(x1.device)
        slot = ring.acquire()
        try:
            loss, _stats = ContrastiveLossFunction.apply(x1, x2, kind, temperature, normalize, None, None, slot)
        except BaseException:
            ring.abandon(slot)
            raise
        correct = ring.wait(slot)[2]
    else:
        loss, stats = ContrastiveLossFunction.apply(x1, x2, kind, temperature, normalize, weight, None)
        correct = stats[2].item()
    return loss, 100.0 * correct / (2 * x1.shape[0])


def contrastive_loss(x_batch1, x_batch2, temperature=1.0, normalize=True, weight=None):
    """NT-Xent loss and auxiliary-task top-1 accuracy (reference objective.py:6-55)."""
    x1, x2 = _as_supported(x_batch1), _as_supported(x_batch2)
    return _loss_and_accuracy(x1, x2, LOSS_NTXENT, float(temperature), bool(normalize), weight)


def modified_contrastive_loss(x_batch1, x_batch2, **kwargs):
    """Probabilistic ("--new_loss" / --modified_loss) variant (reference objective.py:58-98)."""
    temperature = kwargs.get("temperature", 1.0)   # objective.py:68: every other kwarg is ignored
    x1, x2 = _as_supported(x_batch1), _as_supported(x_batch2)
    return _loss_and_accuracy(x1, x2, LOSS_MODIFIED, float(temperature), True, None)
