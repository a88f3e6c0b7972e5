"""B200-native contrastive objective for SimCLR (NT-Xent + the probabilistic "modified" loss).

Import name: ``pytorch_simclr_b200`` (the directory is ``pytorch-simclr_b200/``; the root-level
``pytorch_simclr_b200.py`` shim points Python at it).
"""
from .objective import contrastive_loss, modified_contrastive_loss, set_lazy_accuracy, get_lazy_accuracy  # noqa: F401
from .functional import (ContrastiveLossFunction, contrastive_forward_backward, LOSS_NTXENT,  # noqa: F401
                         LOSS_MODIFIED, set_precision, get_precision, set_eager_backward, get_eager_backward,
                         set_deterministic, get_deterministic)
from .head import bn_contrastive_loss, bn_modified_contrastive_loss  # noqa: F401

__all__ = ["contrastive_loss", "modified_contrastive_loss", "ContrastiveLossFunction",
           "contrastive_forward_backward", "LOSS_NTXENT", "LOSS_MODIFIED", "set_precision", "get_precision",
           "set_eager_backward", "get_eager_backward", "set_deterministic", "get_deterministic", "bn_contrastive_loss",
           "bn_modified_contrastive_loss", "set_lazy_accuracy", "get_lazy_accuracy"]
