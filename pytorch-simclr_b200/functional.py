"""torch.autograd.Function wrappers around the C ABI (single GPU and row-sharded global batch).

The tensors handed to the library are allocated by PyTorch's caching allocator and kept alive by the
autograd context; launches go to ``torch.cuda.current_stream()``.  Nothing here computes the loss in
PyTorch: if the extension is missing or the tensors are not on a B200 the call raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _hoststats, _lib
from ._lib import DTYPE_BF16, DTYPE_F32, LOSS_MODIFIED, LOSS_NTXENT, PRECISION_BF16, PRECISION_SPLIT, check

__all__ = ["contrastive_forward_backward", "ContrastiveLossFunction", "LOSS_NTXENT", "LOSS_MODIFIED", "pad_rows",
           "pad_dim", "compact_to_padded", "padded_to_compact", "set_precision", "get_precision", "resolve_precision",
           "set_eager_backward", "get_eager_backward", "run_fused", "run_fused_begin", "run_fused_finish",
           "set_deterministic", "get_deterministic"]

BLOCK = 128

# Arithmetic of the tensor-core products (include/simclr_b200.h SIMCLR_PRECISION_*):
#   "bf16"  bf16 operands, fp32 accumulate -- the fast path (loss 2e-3, gradients 1e-2 against the fp32 reference)
#   "fp32"  hi + lo bf16 operand planes, three products each -- fp32-grade (loss 1e-5, gradients 1e-4), d <= 128
#   "auto"  fp32-grade for float32 inputs of d <= 128 on one GPU (what the fp32 reference delivers), bf16 otherwise
_PRECISION = "auto"


# When a single-GPU, unweighted loss is evaluated with gradients enabled, the autograd Function runs the first FOUR
# kernels of the fused step at forward time (simclr_forward_backward_begin: the backward tile kernel is already in flight
# when forward() returns) and backward() launches the last one with grad_output as its device scalar
# (simclr_forward_backward_finish).  The reference's call pattern (loss, acc = loss_fn(...); loss /= k; loss.backward(),
# utils/model_utils.py:115-120) always follows a training forward by its backward, and the accuracy read-back it forces
# (objective.py:52) otherwise splits the step into two launch sequences with a host round trip in between.
_EAGER_BACKWARD = True


# Run-to-run reproducible gradients (include/simclr_b200.h SIMCLR_FLAG_DETERMINISTIC): None follows
# torch.are_deterministic_algorithms_enabled() -- the modern form of the reference's cudnn.deterministic switch
# (pretrain.py:59-61) --, True / False force it.  The deterministic backward is a few percent slower.
_DETERMINISTIC = None


def set_deterministic(flag) -> None:
    """True / False, or None to follow torch.use_deterministic_algorithms()."""
    global _DETERMINISTIC
    _DETERMINISTIC = None if flag is None else bool(flag)


def get_deterministic() -> bool:
    return torch.are_deterministic_algorithms_enabled() if _DETERMINISTIC is None else _DETERMINISTIC


def backward_flags() -> int:
    return _lib.FLAG_DETERMINISTIC if get_deterministic() else 0


def set_eager_backward(flag: bool) -> None:
    """True (default): training forwards of the single-GPU unweighted losses compute their gradients at once."""
    global _EAGER_BACKWARD
    _EAGER_BACKWARD = bool(flag)


def get_eager_backward() -> bool:
    return _EAGER_BACKWARD


def set_precision(mode: str) -> None:
    """Select the arithmetic of the loss for subsequent calls: "auto" (default), "bf16" or "fp32"."""
    global _PRECISION
    if mode not in ("auto", "bf16", "fp32"):
        raise ValueError("precision must be 'auto', 'bf16' or 'fp32'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def resolve_precision(x: torch.Tensor, distributed: bool = False, mode: Optional[str] = None) -> int:
    mode = mode or _PRECISION
    if mode == "bf16":
        return PRECISION_BF16
    fits = x.shape[1] <= 128
    if mode == "fp32":
        if not fits:
            raise ValueError("precision 'fp32' (split bf16 operands) supports embedding dimensions up to 128")
        return PRECISION_SPLIT
    return PRECISION_SPLIT if (x.dtype == torch.float32 and fits and not distributed) else PRECISION_BF16


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _device_guard(dev: torch.device):
    """torch.cuda.device(dev) only when dev is not already the current device (the context manager costs microseconds
    on a path whose whole host side is a few tens of them)."""
    idx = dev.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(dev)


def pad_rows(b: int) -> int:
    return (b + BLOCK - 1) // BLOCK * BLOCK


def pad_dim(d: int) -> int:
    if d < 1 or d > 256:
        raise ValueError(f"embedding dimension {d} is not supported (1..256)")
    return 64 if d <= 64 else (128 if d <= 128 else 256)


def compact_to_padded(x: torch.Tensor, b: int) -> torch.Tensor:
    """[2b] in the reference's order (view 1 rows, view 2 rows) -> [2*pad_rows(b)] view-padded, zero filled."""
    bp = pad_rows(b)
    out = x.new_zeros(2 * bp)
    out[:b] = x[:b]
    out[bp:bp + b] = x[b:]
    return out


def padded_to_compact(x: torch.Tensor, b: int) -> torch.Tensor:
    bp = x.shape[0] // 2
    return torch.cat((x[:b], x[bp:bp + b]))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise ValueError(f"unsupported dtype {t.dtype}: pass float32 or bfloat16 embeddings")


def _validate(x1: torch.Tensor, x2: torch.Tensor) -> Tuple[int, int]:
    if x1.dim() != 2 or x1.shape != x2.shape:
        raise ValueError(f"x_batch1 / x_batch2 must both be [batch, dim]; got {tuple(x1.shape)} and {tuple(x2.shape)}")
    if x1.dtype != x2.dtype:
        raise ValueError("x_batch1 and x_batch2 must share a dtype")
    if not (x1.is_cuda and x2.is_cuda) or x1.device != x2.device:
        raise ValueError("simclr_b200 runs on a B200 only: both batches must live on the same CUDA device "
                         "(there is no CPU fallback; the CPU oracle lives in oracle/ for tests)")
    b, d = x1.shape
    if b < 1 or d < 1:
        raise ValueError("empty batch")
    return b, d


class _Saved:
    """State a forward leaves for its backward (device buffers only)."""
    __slots__ = ("operand_rows", "operand_cols", "inv_norm", "pos_dot", "lse2_cols", "col_scale", "b_local",
                 "b_global", "row_offset", "d", "loss", "temperature", "normalize", "dtype_code", "peer", "generation",
                 "bwd_ws", "primed_colvec", "precision")


def run_forward(loss_kind: int, x1: torch.Tensor, x2: torch.Tensor, temperature: float, normalize: bool,
                weight: Optional[torch.Tensor], gather=None, prime_backward: bool = False,
                precision: Optional[int] = None, host_slot: Optional[int] = None):
    """prepare + forward through the C ABI.  ``gather`` (distributed.py) turns the local operand / lse2 into
    their global-batch counterparts and returns (operand_cols, b_global, row_offset, reducer)."""
    lib = _lib.load()
    b, d = _validate(x1, x2)
    x1 = x1.contiguous()
    x2 = x2.contiguous()
    dev = x1.device
    bp, dp = pad_rows(b), pad_dim(d)
    code = _dtype_code(x1)
    if precision is None:
        precision = resolve_precision(x1, gather is not None)
    planes = 2 if precision == PRECISION_SPLIT else 1
    with _device_guard(dev):
        stream = torch.cuda.current_stream().cuda_stream
        operand = torch.empty((planes * 2 * bp, dp), dtype=torch.bfloat16, device=dev)
        rowvec = torch.empty((4, 2 * bp), dtype=torch.float32, device=dev)   # inv_norm, pos_dot, lse2, row_loss
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        b_global_hint = b if gather is None else b * gather.world
        ws_bytes = lib.simclr_forward_workspace_bytes(loss_kind, b, b_global_hint, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        # exact fp32 normalised rows for the accuracy candidates of OTHER ranks (collective transport: gathered below)
        zrows_local = None
        if gather is not None and precision == PRECISION_BF16 and hasattr(gather, "zrows"):
            zrows_local = torch.empty((2 * bp, dp), dtype=torch.float32, device=dev)
        check(lib.simclr_prepare_peer(loss_kind, x1.data_ptr(), x2.data_ptr(), b, d, code, int(bool(normalize)),
                                      float(temperature), precision, operand.data_ptr(), rowvec[0].data_ptr(),
                                      rowvec[1].data_ptr(), ws.data_ptr(), 0, 0, None, None, _ptr(zrows_local), stream),
              "simclr_prepare")
        zrows_global = None
        if gather is None:
            operand_cols, b_global, row_offset = operand, b, 0
        else:
            operand_cols, b_global, row_offset = gather.operand(operand, b)
            if zrows_local is not None:
                zrows_global = gather.zrows(zrows_local, b)
        w_local = None
        if weight is not None:
            w_local = weight.to(device=dev, dtype=torch.float32).contiguous()
            if w_local.numel() != 2 * b:
                raise ValueError(f"weight must have {2 * b} entries, got {w_local.numel()}")
        # single GPU, unweighted, gradients wanted: the forward also prepares the backward workspace (zeroed
        # accumulation buffer + column vectors), which saves the backward-prepare kernel
        bwd_ws = None
        if prime_backward and gather is None and weight is None:
            bwd_bytes = lib.simclr_backward_workspace_bytes_flags(loss_kind, b, b, d, backward_flags())
            bwd_ws = torch.empty(bwd_bytes, dtype=torch.uint8, device=dev)
        stats_ptr = stats.data_ptr()
        if host_slot is not None and gather is None:
            stats_ptr = _hoststats.ring(dev).pointer(host_slot)
            stats = _hoststats.ring(dev).buf[host_slot]
        check(lib.simclr_forward_peer(loss_kind, operand.data_ptr(), operand_cols.data_ptr(), b, b_global, row_offset, d,
                                      float(temperature), int(bool(normalize)), precision, rowvec[1].data_ptr(),
                                      _ptr(w_local),
                                      rowvec[2].data_ptr(), rowvec[3].data_ptr(), stats_ptr, loss.data_ptr(),
                                      ws.data_ptr(), ws_bytes, _ptr(bwd_ws), 0 if bwd_ws is None else bwd_ws.numel(), 0, 0,
                                      None, None, None, None, x1.data_ptr(), x2.data_ptr(), code, rowvec[0].data_ptr(),
                                      _ptr(zrows_global), None, stream), "simclr_forward")
        if host_slot is not None and gather is None:
            _hoststats.ring(dev).launched(host_slot)
    saved = _Saved()
    saved.operand_rows, saved.operand_cols = operand, operand_cols
    saved.inv_norm, saved.pos_dot = rowvec[0], rowvec[1]
    saved.lse2_cols = rowvec[2]
    saved.col_scale = None
    saved.b_local, saved.b_global, saved.row_offset, saved.d = b, b_global, row_offset, d
    saved.loss, saved.temperature, saved.normalize, saved.dtype_code = loss_kind, float(temperature), bool(normalize), code
    saved.bwd_ws, saved.primed_colvec = bwd_ws, (None if bwd_ws is None else bwd_ws.data_ptr())
    saved.precision = precision
    if gather is not None:
        loss, stats = gather.reduce(stats, loss)
        saved.lse2_cols = gather.rowvec(rowvec[2], b)
        if weight is not None:
            saved.col_scale = gather.col_scale(w_local, stats, b)
    elif weight is not None:
        saved.col_scale = compact_to_padded(w_local / w_local.sum(), b)
    return loss, stats, rowvec, saved


_FUSED_PLAN = {}      # (loss, b, d, precision) -> byte offsets of the scratch carved out of one allocation


def _fused_plan(lib, loss_kind: int, b: int, d: int, precision: int, flags: int = 0):
    key = (loss_kind, b, d, precision, flags)
    plan = _FUSED_PLAN.get(key)
    if plan is None:
        bp = pad_rows(b)
        al = lambda n: (n + 255) // 256 * 256      # noqa: E731
        op_bytes = lib.simclr_operand_bytes(b, d, precision)
        if not op_bytes:
            raise ValueError("unsupported shape / precision (fp32-grade operands need d <= 128)")
        fwd_bytes = lib.simclr_forward_workspace_bytes(loss_kind, b, b, d)
        bwd_bytes = lib.simclr_backward_workspace_bytes_flags(loss_kind, b, b, d, flags)
        off, plan = 0, {"flags": flags}
        for name, n in (("operand", op_bytes), ("rowvec", 4 * 2 * bp * 4), ("stats", 16), ("fwd", fwd_bytes),
                        ("bwd", bwd_bytes)):
            plan[name] = (off, n)
            off += al(n)
        plan["total"] = off
        _FUSED_PLAN[key] = plan
    return plan


def run_fused(loss_kind: int, x1: torch.Tensor, x2: torch.Tensor, temperature: float, normalize: bool,
              grad_out: Optional[torch.Tensor] = None, precision: Optional[int] = None, host_slot: Optional[int] = None):
    """The fused five-kernel step through simclr_forward_backward (one GPU, unweighted): returns
    (loss, stats, grad1, grad2), gradients of grad_out * loss (grad_out: device scalar or None for 1).
    One scratch allocation per call (operand, row vectors, statistics and both workspaces are carved out of it)."""
    lib = _lib.load()
    b, d = _validate(x1, x2)
    x1 = x1.contiguous()
    x2 = x2.contiguous()
    dev = x1.device
    code = _dtype_code(x1)
    if precision is None:
        precision = resolve_precision(x1, False)
    plan = _fused_plan(lib, loss_kind, b, d, precision, backward_flags())
    with _device_guard(dev):
        stream = torch.cuda.current_stream().cuda_stream
        scratch = torch.empty(plan["total"], dtype=torch.uint8, device=dev)
        base = scratch.data_ptr()
        loss = torch.empty((), dtype=torch.float32, device=dev)
        g1 = torch.empty_like(x1)
        g2 = torch.empty_like(x2)
        go = None
        if grad_out is not None:
            go = grad_out.to(device=dev, dtype=torch.float32).contiguous()
        if host_slot is not None:
            # the statistics go straight to pinned host memory (see _hoststats.py); `stats` is then that host row
            stats_ptr = _hoststats.ring(dev).pointer(host_slot)
            stats = _hoststats.ring(dev).buf[host_slot]
        else:
            so = plan["stats"][0]
            stats = scratch[so:so + 16].view(torch.float32)
            stats_ptr = base + so
        check(lib.simclr_forward_backward(loss_kind, x1.data_ptr(), x2.data_ptr(), b, d, code, int(bool(normalize)),
                                          float(temperature), precision, _ptr(go), base + plan["operand"][0],
                                          base + plan["rowvec"][0], stats_ptr, loss.data_ptr(), g1.data_ptr(),
                                          g2.data_ptr(), base + plan["fwd"][0], plan["fwd"][1], base + plan["bwd"][0],
                                          plan["bwd"][1], plan["flags"], stream),
              "simclr_forward_backward")
        if host_slot is not None:
            _hoststats.ring(dev).launched(host_slot)
    return loss, stats, g1, g2


class _FusedState:
    """What simclr_forward_backward_begin leaves for simclr_forward_backward_finish (device buffers + call constants)."""
    __slots__ = ("scratch", "plan", "loss", "b", "d", "code", "normalize", "temperature", "precision")


def run_fused_begin(loss_kind: int, x1: torch.Tensor, x2: torch.Tensor, temperature: float, normalize: bool,
                    precision: Optional[int] = None, host_slot: Optional[int] = None):
    """First four kernels of the fused step (simclr_forward_backward_begin): returns (loss, stats, state).  The loss
    statistics are complete when the forward finalize kernel has run -- the backward tile kernel, launched by the same
    call, is still running then; ``run_fused_finish`` launches the last kernel once the upstream gradient is known."""
    lib = _lib.load()
    b, d = _validate(x1, x2)
    dev = x1.device
    if precision is None:
        precision = resolve_precision(x1, False)
    plan = _fused_plan(lib, loss_kind, b, d, precision, backward_flags())
    st = _FusedState()
    st.plan, st.loss, st.b, st.d, st.code = plan, loss_kind, b, d, _dtype_code(x1)
    st.normalize, st.temperature, st.precision = int(bool(normalize)), float(temperature), precision
    with _device_guard(dev):
        stream = torch.cuda.current_stream().cuda_stream
        st.scratch = scratch = torch.empty(plan["total"], dtype=torch.uint8, device=dev)
        base = scratch.data_ptr()
        loss = torch.empty((), dtype=torch.float32, device=dev)
        if host_slot is not None:
            ring = _hoststats.ring(dev)
            stats_ptr, stats = ring.pointer(host_slot), ring.buf[host_slot]
        else:
            so = plan["stats"][0]
            stats = scratch[so:so + 16].view(torch.float32)
            stats_ptr = base + so
        check(lib.simclr_forward_backward_begin(loss_kind, x1.data_ptr(), x2.data_ptr(), b, d, st.code, st.normalize,
                                                st.temperature, precision, base + plan["operand"][0],
                                                base + plan["rowvec"][0], stats_ptr, loss.data_ptr(),
                                                base + plan["fwd"][0], plan["fwd"][1], base + plan["bwd"][0],
                                                plan["bwd"][1], plan["flags"], stream), "simclr_forward_backward_begin")
        if host_slot is not None:
            ring.launched(host_slot)
    return loss, stats, st


def run_fused_finish(st: "_FusedState", x1: torch.Tensor, x2: torch.Tensor, grad_out: Optional[torch.Tensor]):
    """The backward finalize kernel (simclr_forward_backward_finish): gradients of grad_out * loss.  Repeatable."""
    lib = _lib.load()
    dev = x1.device
    plan = st.plan
    with _device_guard(dev):
        stream = torch.cuda.current_stream().cuda_stream
        g1 = torch.empty_like(x1)
        g2 = torch.empty_like(x2)
        go = None
        if grad_out is not None:
            go = grad_out
            if go.dtype != torch.float32 or go.device != dev or not go.is_contiguous():
                go = go.to(device=dev, dtype=torch.float32).contiguous()
        base = st.scratch.data_ptr()
        check(lib.simclr_forward_backward_finish(st.loss, x1.data_ptr(), x2.data_ptr(), st.b, st.d, st.code, st.normalize,
                                                 st.temperature, st.precision, _ptr(go), base + plan["operand"][0],
                                                 base + plan["rowvec"][0], g1.data_ptr(), g2.data_ptr(),
                                                 base + plan["bwd"][0], plan["bwd"][1], plan["flags"], stream),
              "simclr_forward_backward_finish")
    return g1, g2


def run_backward(saved: "_Saved", x1: torch.Tensor, x2: torch.Tensor, grad_out: Optional[torch.Tensor]):
    lib = _lib.load()
    dev = x1.device
    peer = getattr(saved, "peer", None)
    if peer is not None and peer.generation - saved.generation >= peer.GENERATIONS:
        raise RuntimeError("the peer-memory buffers of this forward have been reused: call backward before the "
                           f"{peer.GENERATIONS}nd following global forward (or use transport='nccl')")
    with _device_guard(dev):
        stream = torch.cuda.current_stream().cuda_stream
        g1 = torch.empty_like(x1)
        g2 = torch.empty_like(x2)
        go = None
        if grad_out is not None:
            go = grad_out.to(device=dev, dtype=torch.float32).contiguous()
        ws = getattr(saved, "bwd_ws", None)
        primed = getattr(saved, "primed_colvec", None) if ws is not None else None
        flags = backward_flags()
        need = lib.simclr_backward_workspace_bytes_flags(saved.loss, saved.b_local, saved.b_global, saved.d, flags)
        if ws is not None and ws.numel() < need:
            ws, primed = None, None       # the switch was flipped between forward and backward: take the unprimed route
        if ws is None:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
        ws_bytes = ws.numel()
        check(lib.simclr_backward(saved.loss, x1.data_ptr(), x2.data_ptr(), saved.b_local, saved.b_global,
                                  saved.row_offset, saved.d, saved.dtype_code, int(saved.normalize), saved.temperature,
                                  getattr(saved, "precision", PRECISION_BF16), saved.operand_rows.data_ptr(), saved.operand_cols.data_ptr(),
                                  saved.inv_norm.data_ptr(), saved.pos_dot.data_ptr(), saved.lse2_cols.data_ptr(),
                                  _ptr(saved.col_scale), _ptr(go), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(),
                                  ws_bytes, primed, flags, stream), "simclr_backward")
        # the primed state (zeroed accumulation buffer) is consumed: a second backward over a retained graph takes
        # the unprimed route (its backward-prepare kernel zeroes the buffer again)
        saved.primed_colvec = None
    return g1, g2


class ContrastiveLossFunction(torch.autograd.Function):
    """(x_batch1, x_batch2) -> (loss, stats).  stats = [sum w L, sum w, #correct, loss] is not differentiable.

    The returned ``loss`` is its own 0-d tensor (not a view, not saved), so the caller may divide it in place
    (reference utils/model_utils.py:31,116); backward honours the resulting ``grad_output``.
    """

    @staticmethod
    def forward(ctx, x1, x2, loss_kind, temperature, normalize, weight, gather, host_slot=None):
        if getattr(gather, "peer", False):
            from .distributed import run_forward_peer
            loss, stats, _rowvec, saved = run_forward_peer(loss_kind, x1, x2, temperature, normalize, gather,
                                                           any(ctx.needs_input_grad[:2]))
        else:
            want_grad = any(ctx.needs_input_grad[:2])
            if want_grad and _EAGER_BACKWARD and gather is None and weight is None:
                # Four of the five kernels of the fused step now (the loss / accuracy are complete after the third, so the
                # caller reads them while the backward tile kernel runs); backward() launches the finalize kernel with
                # grad_output as its device scalar -- no extra pass over the gradients.
                x1, x2 = x1.contiguous(), x2.contiguous()
                loss, stats, state = run_fused_begin(loss_kind, x1, x2, temperature, normalize, host_slot=host_slot)
                ctx.fused_state = state
                ctx.save_for_backward(x1, x2)
                ctx.mark_non_differentiable(stats)
                return loss, stats
            loss, stats, _rowvec, saved = run_forward(loss_kind, x1, x2, temperature, normalize, weight, gather, want_grad,
                                                      host_slot=host_slot if weight is None else None)
        ctx.fused_state = None
        ctx.saved_state = saved
        ctx.save_for_backward(x1, x2)
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        x1, x2 = ctx.saved_tensors
        if ctx.fused_state is not None:
            g1, g2 = run_fused_finish(ctx.fused_state, x1, x2, grad_loss)
            return g1, g2, None, None, None, None, None, None
        g1, g2 = run_backward(ctx.saved_state, x1.contiguous(), x2.contiguous(), grad_loss)
        return g1, g2, None, None, None, None, None, None


def contrastive_forward_backward(loss_kind: int, x1: torch.Tensor, x2: torch.Tensor, temperature: float,
                                 normalize: bool = True, weight: Optional[torch.Tensor] = None,
                                 grad_out: Optional[torch.Tensor] = None, precision: Optional[str] = None):
    """Autograd-free fused call: returns (loss, stats, grad1, grad2), all on the device."""
    prec = None if precision is None else resolve_precision(x1, False, precision)
    if weight is None:
        return run_fused(loss_kind, x1, x2, temperature, normalize, grad_out, prec)
    loss, stats, _rowvec, saved = run_forward(loss_kind, x1, x2, temperature, normalize, weight, None, True, prec)
    g1, g2 = run_backward(saved, x1.contiguous(), x2.contiguous(), grad_out)
    return loss, stats, g1, g2
