"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/simclr_b200.h declares,
and rejects bad arguments with the documented error codes before touching a GPU."""
import ctypes
import os
import re

import pytest

from conftest import REPO

HEADER = os.path.join(REPO, "include", "simclr_b200.h")
DEBUG_HEADER = os.path.join(REPO, "include", "simclr_b200_debug.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    from pytorch_simclr_b200 import _lib
    return _lib.load()


def _declared_functions(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(simclr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from pytorch_simclr_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 11
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/simclr_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_diagnostics_live_in_the_tracing_build_only(lib):
    """include/simclr_b200_debug.h: exported by libsimclr_b200_trace.so (next to the whole product ABI), absent from
    the product library."""
    from pytorch_simclr_b200 import _lib
    debug = _declared_functions(DEBUG_HEADER)
    assert sorted(_lib.DEBUG_SIGNATURES) == debug
    product = ctypes.CDLL(_lib.LIB_PATH)
    trace = ctypes.CDLL(_lib.TRACE_LIB_PATH)
    for name in debug:
        assert not hasattr(product, name), f"{name} must not ship in the product library"
        assert hasattr(trace, name)
    for name in _declared_functions():
        assert hasattr(trace, name)


def test_abi_version_and_error_strings(lib):
    assert lib.simclr_abi_version() == 13
    assert lib.simclr_error_string(0) == b"ok"
    for code in range(-11, 0):
        assert lib.simclr_error_string(code) not in (b"", b"unknown error")
    assert lib.simclr_error_string(-99) == b"unknown error"


def test_padding_helpers(lib):
    assert [lib.simclr_pad_rows(b) for b in (1, 127, 128, 129, 4096)] == [128, 128, 128, 256, 4096]
    assert [lib.simclr_pad_dim(d) for d in (1, 64, 65, 128, 129, 256)] == [64, 64, 128, 128, 256, 256]
    assert lib.simclr_pad_dim(257) == 0 and lib.simclr_pad_dim(0) == 0 and lib.simclr_pad_rows(0) == 0


def test_workspace_sizes(lib):
    for loss in (0, 1):
        f = lib.simclr_forward_workspace_bytes(loss, 4096, 4096, 128)
        b = lib.simclr_backward_workspace_bytes(loss, 4096, 4096, 128)
        assert f > 0 and b >= 2 * 4096 * 128 * 4
    assert lib.simclr_forward_workspace_bytes(0, 0, 0, 128) == 0          # bad shape
    assert lib.simclr_forward_workspace_bytes(0, 64, 64, 300) == 0        # unsupported dim
    assert lib.simclr_forward_workspace_bytes(7, 64, 64, 128) == 0        # unknown loss
    # the deterministic backward keeps one accumulator slot per (CTA, segment)
    assert lib.simclr_backward_workspace_bytes_flags(0, 4096, 4096, 128, 1) > lib.simclr_backward_workspace_bytes(0, 4096, 4096, 128)
    assert lib.simclr_backward_workspace_bytes_flags(0, 4096, 4096, 128, 0) == lib.simclr_backward_workspace_bytes(0, 4096, 4096, 128)
    # sharded rows need less accumulator space than the whole batch
    assert lib.simclr_backward_workspace_bytes(0, 512, 4096, 128) < lib.simclr_backward_workspace_bytes(0, 4096, 4096, 128)


def test_argument_validation_without_gpu(lib):
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.addressof(buf)
    p = (p + 255) & ~255
    # null pointers
    assert lib.simclr_prepare(0, None, p, 4, 8, 0, 1, 0.5, p, p, p, None, None) == -1
    assert lib.simclr_forward(0, None, p, 4, 4, 0, 8, 0.5, 1, p, None, p, p, p, None, p, 1 << 15, None, None, 0, None, None) == -1
    # bad dtype / shape / dim / loss / temperature / workspace
    assert lib.simclr_prepare(0, p, p, 4, 8, 9, 1, 0.5, p, p, p, None, None) == -4
    assert lib.simclr_prepare(0, p, p, 0, 8, 0, 1, 0.5, p, p, p, None, None) == -2
    assert lib.simclr_prepare(0, p, p, 4, 300, 0, 1, 0.5, p, p, p, None, None) == -3
    assert lib.simclr_prepare(5, p, p, 4, 8, 0, 1, 0.5, p, p, p, None, None) == -11
    assert lib.simclr_forward(0, p, p, 4, 4, 0, 8, 0.0, 1, p, None, p, p, p, None, p, 1 << 15, None, None, 0, None, None) == -7
    assert lib.simclr_forward(0, p, p, 4, 2, 0, 8, 0.5, 1, p, None, p, p, p, None, p, 1 << 15, None, None, 0, None, None) == -2   # b_glob < b_loc
    assert lib.simclr_forward(0, p, p, 4, 8, 6, 8, 0.5, 1, p, None, p, p, p, None, p, 1 << 15, None, None, 0, None, None) == -2   # shard outside
    assert lib.simclr_forward(0, p, p, 4, 4, 0, 8, 0.5, 1, p, None, p, p, p, None, p, 16, None, None, 0, None, None) == -5
    assert lib.simclr_forward(0, p + 4, p, 4, 4, 0, 8, 0.5, 1, p, None, p, p, p, None, p, 1 << 15, None, None, 0, None, None) == -6
    assert lib.simclr_backward(0, p, p, 4, 4, 0, 8, 0, 1, float("nan"), 0, p, p, p, p, p, None, None, p, p, p, 1 << 15,
                               None, 0, None) == -7
