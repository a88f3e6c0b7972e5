"""Projection-head tail fused into the loss.

The reference's projection head ends in ``Linear(encoder_dim, out_dim, bias=False)`` + ``BatchNorm1d(out_dim)``
(reference models/simclr.py:38-39) and the training loop calls the model once per view (utils/model_utils.py:113-114),
so each view is normalised with its own batch statistics, before ``loss_fn(z1, z2, temperature=...)`` (:115).

    loss, acc = bn_contrastive_loss(u1, u2, bn, temperature)        # == contrastive_loss(bn(u1), bn(u2), temperature)

takes the PRE-BatchNorm activations ``u = g_linear(...)`` and the module itself: the BatchNorm apply runs in the registers of
the loss's prepare kernel (``z`` is never written to memory), the backward returns dL/du and accumulates dL/dgamma, dL/dbeta,
and the module's running statistics are updated exactly as two successive ``bn(u1)``, ``bn(u2)`` calls would (training mode);
in eval mode the running statistics are used.  One GPU, unweighted losses; CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _hoststats, _lib
from ._lib import check
from .functional import (LOSS_MODIFIED, LOSS_NTXENT, _device_guard, _dtype_code, _fused_plan, _ptr, _validate, backward_flags,
                         pad_dim, resolve_precision)

__all__ = ["bn_contrastive_loss", "bn_modified_contrastive_loss", "HeadTailFunction"]


class _BN(ctypes.Structure):
    """simclr_bn_t of include/simclr_b200.h."""
    _fields_ = [("state", ctypes.c_void_p), ("partial", ctypes.c_void_p)]


_WS = {}      # (device, b, d) -> zero-initialised workspace (its ticket header is left zero by every call)


def _workspace(lib, dev, b: int, d: int) -> torch.Tensor:
    key = (dev, b, d)
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.zeros(lib.simclr_bn_workspace_bytes(b, d), dtype=torch.uint8, device=dev)
    return ws


def _eval_state(bn: torch.nn.BatchNorm1d, d: int, dp: int, dev) -> torch.Tensor:
    """scale / shift from the running statistics (the same for both views); planes 2-4 are not used in eval mode."""
    st = torch.zeros((2, 5, dp), dtype=torch.float32, device=dev)
    gamma = bn.weight.detach().float() if bn.affine else torch.ones(d, device=dev)
    beta = bn.bias.detach().float() if bn.affine else torch.zeros(d, device=dev)
    scale = gamma / torch.sqrt(bn.running_var.float() + bn.eps)
    st[:, 0, :d] = scale
    st[:, 1, :d] = beta - bn.running_mean.float() * scale
    return st


class HeadTailFunction(torch.autograd.Function):
    """(u1, u2, gamma, beta) -> (loss, stats, bn_state); bn_state f32 [2][5][Dpad] (see include/simclr_b200.h)."""

    @staticmethod
    def forward(ctx, u1, u2, gamma, beta, loss_kind, temperature, normalize, eps, training, eval_state, host_slot):
        lib = _lib.load()
        b, d = _validate(u1, u2)
        u1, u2 = u1.contiguous(), u2.contiguous()
        dev = u1.device
        dp = pad_dim(d)
        code = _dtype_code(u1)
        precision = resolve_precision(u1, False)
        flags = backward_flags()
        plan = _fused_plan(lib, loss_kind, b, d, precision, flags)
        want_grad = any(ctx.needs_input_grad[:4])
        with _device_guard(dev):
            stream = torch.cuda.current_stream().cuda_stream
            ws = _workspace(lib, dev, b, d)
            if training:
                state = torch.empty((2, 5, dp), dtype=torch.float32, device=dev)
                g32 = gamma.detach().float().contiguous()
                b32 = beta.detach().float().contiguous()
                check(lib.simclr_bn_stats(u1.data_ptr(), u2.data_ptr(), b, d, code, g32.data_ptr(), b32.data_ptr(), float(eps),
                                          state.data_ptr(), ws.data_ptr(), ws.numel(), stream), "simclr_bn_stats")
            else:
                state = eval_state
            scratch = torch.empty(plan["total"], dtype=torch.uint8, device=dev)
            base = scratch.data_ptr()
            loss = torch.empty((), dtype=torch.float32, device=dev)
            if host_slot is not None:
                ring = _hoststats.ring(dev)
                stats_ptr, stats = ring.pointer(host_slot), ring.buf[host_slot]
            else:
                so = plan["stats"][0]
                stats = scratch[so:so + 16].view(torch.float32)
                stats_ptr = base + so
            bn = _BN(state.data_ptr(), ws.data_ptr() + 256)
            if want_grad:
                check(lib.simclr_head_forward_backward_begin(
                    loss_kind, u1.data_ptr(), u2.data_ptr(), b, d, code, int(bool(normalize)), float(temperature), precision,
                    ctypes.byref(bn), base + plan["operand"][0], base + plan["rowvec"][0], stats_ptr, loss.data_ptr(),
                    base + plan["fwd"][0], plan["fwd"][1], base + plan["bwd"][0], plan["bwd"][1], flags, stream),
                    "simclr_head_forward_backward_begin")
            else:
                check(lib.simclr_head_forward(
                    loss_kind, u1.data_ptr(), u2.data_ptr(), b, d, code, int(bool(normalize)), float(temperature), precision,
                    state.data_ptr(), base + plan["operand"][0], base + plan["rowvec"][0], stats_ptr, loss.data_ptr(),
                    base + plan["fwd"][0], plan["fwd"][1], stream), "simclr_head_forward")
            if host_slot is not None:
                ring.launched(host_slot)
        ctx.save_for_backward(u1, u2)
        ctx.held = (scratch, state, ws, plan, gamma)
        ctx.consts = (loss_kind, b, d, code, int(bool(normalize)), float(temperature), precision, flags, bool(training))
        ctx.mark_non_differentiable(stats, state)
        return loss, stats, state

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats, _grad_state):
        lib = _lib.load()
        u1, u2 = ctx.saved_tensors
        scratch, state, ws, plan, gamma = ctx.held
        loss_kind, b, d, code, normalize, temperature, precision, flags, training = ctx.consts
        dev = u1.device
        with _device_guard(dev):
            stream = torch.cuda.current_stream().cuda_stream
            g1, g2 = torch.empty_like(u1), torch.empty_like(u2)
            ggamma = torch.zeros(d, dtype=torch.float32, device=dev)
            gbeta = torch.zeros(d, dtype=torch.float32, device=dev)
            go = grad_loss
            if go.dtype != torch.float32 or go.device != dev or not go.is_contiguous():
                go = go.to(device=dev, dtype=torch.float32).contiguous()
            base = scratch.data_ptr()
            bn = _BN(state.data_ptr(), ws.data_ptr() + 256)
            check(lib.simclr_head_forward_backward_finish(
                loss_kind, u1.data_ptr(), u2.data_ptr(), b, d, code, normalize, temperature, precision, go.data_ptr(),
                ctypes.byref(bn), int(training), base + plan["operand"][0], base + plan["rowvec"][0], g1.data_ptr(),
                g2.data_ptr(), ggamma.data_ptr(), gbeta.data_ptr(), base + plan["bwd"][0], plan["bwd"][1], flags, stream),
                "simclr_head_forward_backward_finish")
        gg = ggamma.to(gamma.dtype) if ctx.needs_input_grad[2] else None
        gb = gbeta.to(gamma.dtype) if ctx.needs_input_grad[3] else None
        return g1, g2, gg, gb, None, None, None, None, None, None, None


def _run(kind, u1, u2, bn, temperature, normalize):
    if not isinstance(bn, torch.nn.BatchNorm1d):
        raise TypeError("bn must be the nn.BatchNorm1d that closes the projection head (reference models/simclr.py:39)")
    if u1.dtype not in (torch.float32, torch.bfloat16):
        u1, u2 = u1.float(), u2.float()
    b, d = _validate(u1, u2)
    if bn.num_features != d:
        raise ValueError(f"BatchNorm1d has {bn.num_features} features, the activations {d}")
    dev = u1.device
    training = bn.training or bn.running_mean is None
    if training and b < 2:
        raise ValueError("Expected more than 1 value per channel when training (as nn.BatchNorm1d)")
    gamma = bn.weight if bn.affine else torch.ones(d, device=dev)
    beta = bn.bias if bn.affine else torch.zeros(d, device=dev)
    eval_state = None if training else _eval_state(bn, d, pad_dim(d), dev)
    ring = _hoststats.ring(dev)
    slot = ring.acquire()
    try:
        loss, _stats, state = HeadTailFunction.apply(u1, u2, gamma, beta, kind, float(temperature), bool(normalize), bn.eps,
                                                     training, eval_state, slot)
    except BaseException:
        ring.abandon(slot)
        raise
    if training and bn.track_running_stats and bn.running_mean is not None:
        # two successive module calls, view 1 then view 2 (utils/model_utils.py:113-114)
        with torch.no_grad():
            for v in range(2):
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1.0 - m).add_(state[v, 2, :d].to(bn.running_mean.dtype), alpha=m)
                bn.running_var.mul_(1.0 - m).add_(state[v, 4, :d].to(bn.running_var.dtype), alpha=m)
    correct = ring.wait(slot)[2]
    return loss, 100.0 * correct / (2 * b)


def bn_contrastive_loss(u1: torch.Tensor, u2: torch.Tensor, bn: torch.nn.BatchNorm1d, temperature: float = 1.0,
                        normalize: bool = True):
    """``contrastive_loss(bn(u1), bn(u2), temperature, normalize)`` (reference objective.py:6-55 behind models/simclr.py:39)
    without materialising ``bn(u)``; returns ``(loss, acc)`` like the reference."""
    return _run(LOSS_NTXENT, u1, u2, bn, temperature, normalize)


def bn_modified_contrastive_loss(u1: torch.Tensor, u2: torch.Tensor, bn: torch.nn.BatchNorm1d, **kwargs):
    """``modified_contrastive_loss(bn(u1), bn(u2), **kwargs)`` (reference objective.py:58-98)."""
    return _run(LOSS_MODIFIED, u1, u2, bn, kwargs.get("temperature", 1.0), True)
