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


def test_fused_call_scratch_plan_is_aligned_and_sized_by_the_library():
    """functional._fused_plan carves operand, row vectors, statistics and both workspaces out of one allocation."""
    from pytorch_simclr_b200 import _lib
    lib = _lib.load()
    for kind, b, d, prec in ((0, 4096, 128, _lib.PRECISION_BF16), (1, 300, 100, _lib.PRECISION_SPLIT), (0, 777, 256, _lib.PRECISION_BF16)):
        plan = F._fused_plan(lib, kind, b, d, prec)
        names = ["operand", "rowvec", "stats", "fwd", "bwd"]
        end = 0
        for n in names:
            off, size = plan[n]
            assert off % 256 == 0 and off >= end and size > 0
            end = off + size
        assert plan["total"] >= end
        assert plan["operand"][1] == lib.simclr_operand_bytes(b, d, prec)
        assert plan["fwd"][1] == lib.simclr_forward_workspace_bytes(kind, b, b, d)
        assert plan["bwd"][1] == lib.simclr_backward_workspace_bytes(kind, b, b, d)
        assert plan["rowvec"][1] == 4 * 2 * F.pad_rows(b) * 4
    with pytest.raises(ValueError):
        F._fused_plan(lib, 0, 64, 256, _lib.PRECISION_SPLIT)          # split operands need d <= 128


def test_eager_backward_and_precision_switches():
    assert sb.get_eager_backward() is True
    sb.set_eager_backward(False)
    assert sb.get_eager_backward() is False
    sb.set_eager_backward(True)
    with pytest.raises(ValueError):
        sb.set_precision("fp8")
    x = torch.zeros(4, 128)
    assert F.resolve_precision(x, False, "bf16") == 0 and F.resolve_precision(x, False, "fp32") == 1
    assert F.resolve_precision(x, True, "auto") == 0                   # the sharded batch runs bf16 operands
    with pytest.raises(ValueError):
        F.resolve_precision(torch.zeros(4, 200), False, "fp32")


def test_fused_entry_points_reject_bad_arguments_without_a_gpu():
    """simclr_forward_backward / _peer validate before they touch the device."""
    import ctypes
    from pytorch_simclr_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(1 << 12)
    p = (ctypes.addressof(buf) + 255) & ~255
    assert lib.simclr_forward_backward(0, p, p, 4, 8, 0, 1, 0.5, 0, None, p, None, p, None, p, p, p, 1 << 20, p, 1 << 20, 0, None) == -1
    assert lib.simclr_forward_backward(0, p, p, 0, 8, 0, 1, 0.5, 0, None, p, p, p, None, p, p, p, 1 << 20, p, 1 << 20, 0, None) == -2
    arr = (ctypes.c_void_p * 2)(p, p)
    assert lib.simclr_forward_backward_peer(0, p, p, 4, 8, 0, 1, 0.5, None, p, p, p, p, None, p, p, p, 1 << 20, p, 1 << 20, 2, 5,
                                            arr, None, arr, arr, arr, p, None, 0, None) == -12      # rank outside the world
    assert lib.simclr_forward_backward_peer(0, p, p, 4, 8, 0, 1, 0.5, None, p, p, p, p, None, p, p, p, 1 << 20, p, 1 << 20, 2, 0,
                                            None, None, arr, arr, arr, p, None, 0, None) == -1


def test_deterministic_and_lazy_accuracy_switches():
    from pytorch_simclr_b200 import _lib
    assert sb.get_deterministic() is False and F.backward_flags() == 0
    sb.set_deterministic(True)
    try:
        assert sb.get_deterministic() is True and F.backward_flags() == _lib.FLAG_DETERMINISTIC
        lib = _lib.load()
        # the deterministic backward keeps one accumulator slot per (CTA, segment): a larger workspace, sized by the library
        assert lib.simclr_backward_workspace_bytes_flags(0, 4096, 4096, 128, _lib.FLAG_DETERMINISTIC) > \
            lib.simclr_backward_workspace_bytes(0, 4096, 4096, 128)
        assert F._fused_plan(lib, 0, 4096, 128, _lib.PRECISION_BF16, _lib.FLAG_DETERMINISTIC)["bwd"][1] == \
            lib.simclr_backward_workspace_bytes_flags(0, 4096, 4096, 128, _lib.FLAG_DETERMINISTIC)
    finally:
        sb.set_deterministic(False)
    assert sb.get_lazy_accuracy() is False
    sb.set_lazy_accuracy(True)
    assert sb.get_lazy_accuracy() is True
    sb.set_lazy_accuracy(False)


def test_peer_transport_refuses_what_it_cannot_compute():
    """ADVICE (round 1): transport='peer' must not silently drop per-row weights or the fp32-grade mode."""
    from pytorch_simclr_b200 import distributed as D
    z = torch.randn(8, 16)
    with pytest.raises(ValueError, match="per-row weights"):
        D.global_contrastive_loss(z, z, weight=torch.ones(16), transport="peer")
    with pytest.raises(ValueError, match="CUDA"):
        D.global_contrastive_loss(z, z, transport="peer")
    with pytest.raises(ValueError, match="transport must be"):
        D.global_contrastive_loss(z, z, transport="smoke-signals")
    assert "ddp_scale" in inspect.signature(D.global_contrastive_loss).parameters
    assert "ddp_scale" in inspect.signature(D.global_modified_contrastive_loss).parameters


def test_head_tail_argument_checks_without_a_gpu():
    u = torch.randn(8, 16)
    with pytest.raises(TypeError, match="BatchNorm1d"):
        sb.bn_contrastive_loss(u, u, torch.nn.LayerNorm(16))
    with pytest.raises(ValueError, match="no CPU fallback"):
        sb.bn_contrastive_loss(u, u, torch.nn.BatchNorm1d(16))
    lib = __import__("pytorch_simclr_b200._lib", fromlist=["load"]).load()
    assert lib.simclr_bn_state_floats(128) == 2 * 5 * 128 and lib.simclr_bn_workspace_bytes(4096, 128) > 256


def test_library_stamp_ignores_comments_but_not_code(tmp_path, monkeypatch):
    """pytorch-simclr_b200/build.py: profiles are keyed to the stamp; a comment edit must not orphan them."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_build_for_test", os.path.join(root, "pytorch-simclr_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    monkeypatch.setattr(build, "CSRC", str(tmp_path))
    monkeypatch.setattr(build, "DEPS", ["k.cu"])
    src = tmp_path / "k.cu"
    src.write_text("// a kernel\n__global__ void k(int* p) { *p = 1; /* one */ }\n")
    first = build._digest()
    src.write_text("// the same kernel, other words\n\n__global__ void k(int* p) {\n    *p = 1;   /* still one */\n}\n")
    assert build._digest() == first
    src.write_text("// a kernel\n__global__ void k(int* p) { *p = 2; /* one */ }\n")
    assert build._digest() != first
