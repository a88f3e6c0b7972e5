"""Drop-in for the reference's top-level ``objective`` module: put this repository ahead of the reference
on ``sys.path`` and ``from objective import contrastive_loss, modified_contrastive_loss``
(reference utils/model_utils.py:2) resolves to the B200 implementation."""
from pytorch_simclr_b200.objective import contrastive_loss, modified_contrastive_loss  # noqa: F401

__all__ = ["contrastive_loss", "modified_contrastive_loss"]
