"""ctypes binding of include/simclr_b200.h.  There is no CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIMCLR_B200_LIB selects another build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("SIMCLR_B200_LIB") or os.path.join(_HERE, "lib", "libsimclr_b200.so")

LOSS_NTXENT = 0
LOSS_MODIFIED = 1
DTYPE_F32 = 0
DTYPE_BF16 = 1
PRECISION_BF16 = 0
PRECISION_SPLIT = 1
FLAG_DETERMINISTIC = 1
ABI_VERSION = 13

_lock = threading.Lock()
_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/simclr_b200.h declares
SIGNATURES = {
    "simclr_abi_version": (_int, []),
    "simclr_error_string": (ctypes.c_char_p, [_int]),
    "simclr_pad_rows": (_i64, [_i64]),
    "simclr_pad_dim": (_i64, [_i64]),
    "simclr_forward_workspace_bytes": (_sz, [_int, _i64, _i64, _i64]),
    "simclr_backward_workspace_bytes": (_sz, [_int, _i64, _i64, _i64]),
    "simclr_backward_workspace_bytes_flags": (_sz, [_int, _i64, _i64, _i64, _int]),
    "simclr_prepare": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _vp, _vp, _vp, _vp, _vp]),
    "simclr_forward": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _sz, _vp, _vp, _int, _vp, _vp]),
    "simclr_operand_bytes": (_sz, [_i64, _i64, _int]),
    "simclr_backward": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _sz, _vp, _int, _vp]),
    "simclr_forward_backward": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _sz, _vp, _sz, _int, _vp]),
    "simclr_forward_backward_begin": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp,
                                             _sz, _vp, _sz, _int, _vp]),
    "simclr_forward_backward_finish": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp,
                                              _vp, _sz, _int, _vp]),
    "simclr_forward_backward_peer": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32,           # loss .. temperature
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,                  # grad_out .. grad2
                                            _vp, _sz, _vp, _sz, _int, _int,                          # workspaces, world, rank
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp,                       # peers .. zrows_peers
                                            _int, _vp]),
    "simclr_prepare_peer": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _int, _int, _vp,
                                   _vp, _vp, _vp]),
    "simclr_forward_peer": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _sz, _vp, _sz, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "simclr_peer_barrier": (_int, [_int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "simclr_bn_state_floats": (_sz, [_i64]),
    "simclr_bn_workspace_bytes": (_sz, [_i64, _i64]),
    "simclr_bn_stats": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _f32, _vp, _vp, _sz, _vp]),
    "simclr_head_forward": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "simclr_head_forward_backward_begin": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp,
                                                  _vp, _sz, _vp, _sz, _int, _vp]),
    "simclr_head_forward_backward_finish": (_int, [_int, _vp, _vp, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _int, _vp, _vp,
                                                   _vp, _vp, _vp, _vp, _vp, _sz, _int, _vp]),
    "simclr_forward_stages": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _sz, _vp, _sz, _vp, _vp, _int, _vp, _vp, ctypes.c_uint]),
    "simclr_backward_stages": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _sz, _vp, _int, _vp, ctypes.c_uint]),
}

# include/simclr_b200_debug.h: exported by the tracing build (lib/libsimclr_b200_trace.so) only
DEBUG_SIGNATURES = {
    "simclr_selftest_umma": (_int, [_vp, _vp, _vp, _vp]),
    "simclr_debug_set_trace": (_int, [_vp, _int]),
    "simclr_debug_set_kernel_trace": (_int, [_vp]),
    "simclr_debug_chunk_rate": (_int, [_vp, _int, _int, _int, _f32, _vp, _vp]),
    "simclr_debug_pipe_rate": (_int, [_vp, _int, _int, _int, _vp, _vp]),
    "simclr_debug_mma_rate": (_int, [_vp, _int, _int, _int, _vp, _vp]),
}
TRACE_LIB_PATH = os.path.join(_HERE, "lib", "libsimclr_b200_trace.so")
STAGE_PREPARE, STAGE_FORWARD_TILE, STAGE_FORWARD_FINALIZE, STAGE_BACKWARD_TILE, STAGE_BACKWARD_FINALIZE = 1, 2, 4, 8, 16
STAGE_BACKWARD_PREPARE, STAGE_ALL = 32, 0xFFFFFFFF
_debug_lib = None


class SimclrLibraryError(RuntimeError):
    pass


def load():
    """Load libsimclr_b200.so (built by build.py / __graft_entry__.build()); raise loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise SimclrLibraryError(
                f"{LIB_PATH} is missing: build the sm_100a extension first (python pytorch-simclr_b200/build.py). "
                "This package has no CPU or eager fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError -> the library is stale
            fn.restype = res
            fn.argtypes = args
        if lib.simclr_abi_version() != ABI_VERSION:
            raise SimclrLibraryError(f"ABI mismatch: library {lib.simclr_abi_version()} vs binding {ABI_VERSION}")
        _bind_debug(lib)
        _lib = lib
    return _lib


def _bind_debug(lib) -> bool:
    if not hasattr(lib, "simclr_debug_set_trace"):
        return False
    for name, (res, args) in DEBUG_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return True


def load_debug():
    """The tracing build (diagnostics of include/simclr_b200_debug.h next to the whole product ABI).  The timeline tools
    set SIMCLR_B200_LIB to it before importing the package, so that the kernels they trace come from the same library;
    the rate probes and the primitive self-test are self-contained and may load it next to the product library."""
    global _debug_lib
    if _debug_lib is not None:
        return _debug_lib
    lib = load()
    if hasattr(lib, "simclr_debug_set_trace"):
        _debug_lib = lib
        return lib
    with _lock:
        if _debug_lib is None:
            if not os.path.isfile(TRACE_LIB_PATH):
                raise SimclrLibraryError(f"{TRACE_LIB_PATH} is missing: python pytorch-simclr_b200/build.py builds it")
            dbg = ctypes.CDLL(TRACE_LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(dbg, name)
                fn.restype = res
                fn.argtypes = args
            _bind_debug(dbg)
            _debug_lib = dbg
    return _debug_lib


def check(code: int, what: str) -> None:
    """Translate a C-ABI return code (include/simclr_b200.h error convention) into an exception."""
    if code == 0:
        return
    msg = load().simclr_error_string(code).decode()
    if code < 0:
        raise ValueError(f"{what}: {msg} (code {code})")
    raise RuntimeError(f"{what}: CUDA error {code}: {msg}")
