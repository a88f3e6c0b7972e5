"""CPU tests of the host-side mirror of the reference interface (no kernels run here)."""
import inspect

import pytest
import torch

import pytorch_simclr_b200 as sb
from pytorch_simclr_b200 import functional as F


def test_signatures_match_reference_contract():
    """reference objective.py:6-10 and :58-60."""
    sig = inspect.signature(sb.contrastive_loss)
    assert list(sig.parameters) == ["x_batch1", "x_batch2", "temperature", "normalize", "weight"]
    assert sig.parameters["temperature"].default == 1.0
    assert sig.parameters["normalize"].default is True
    assert sig.parameters["weight"].default is None
    sig = inspect.signature(sb.modified_contrastive_loss)
    assert list(sig.parameters) == ["x_batch1", "x_batch2", "kwargs"]
    assert sig.parameters["kwargs"].kind is inspect.Parameter.VAR_KEYWORD


def test_signatures_match_live_reference(reference_objective):
    for name in ("contrastive_loss", "modified_contrastive_loss"):
        assert str(inspect.signature(getattr(sb, name))) == str(inspect.signature(getattr(reference_objective, name)))


def test_top_level_objective_module_is_the_drop_in():
    import objective
    assert objective.contrastive_loss is sb.contrastive_loss
    assert objective.modified_contrastive_loss is sb.modified_contrastive_loss


def test_cpu_tensors_are_rejected_not_silently_computed():
    z = torch.randn(8, 16)
    with pytest.raises(ValueError, match="no CPU fallback"):
        sb.contrastive_loss(z, z)
    with pytest.raises(ValueError, match="no CPU fallback"):
        sb.modified_contrastive_loss(z, z, temperature=0.5)


def test_shape_and_dtype_validation():
    with pytest.raises(ValueError):
        F._validate(torch.randn(4, 8), torch.randn(5, 8))
    with pytest.raises(ValueError):
        F._validate(torch.randn(4, 8), torch.randn(4, 8, dtype=torch.float64))
    with pytest.raises(ValueError):
        F._dtype_code(torch.zeros(1, dtype=torch.float16))
    with pytest.raises(ValueError):
        F.pad_dim(257)


def test_view_padded_layout_round_trip():
    for b in (1, 5, 128, 200):
        x = torch.arange(2 * b, dtype=torch.float32) + 1
        p = F.compact_to_padded(x, b)
        bp = F.pad_rows(b)
        assert p.shape == (2 * bp,)
        assert torch.equal(p[:b], x[:b]) and torch.equal(p[bp:bp + b], x[b:])
        assert p[b:bp].abs().sum() == 0 and p[bp + b:].abs().sum() == 0
        assert torch.equal(F.padded_to_compact(p, b), x)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pytorch_simclr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SimclrLibraryError, match="no CPU or eager fallback"):
        _lib.load()
