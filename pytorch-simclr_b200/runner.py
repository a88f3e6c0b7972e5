"""Allocation-free forward(+backward) runner over the staged C ABI with pre-allocated buffers.

``ContrastiveStep`` is what a training loop that wants CUDA-graph capture (or bench.py) uses: all scratch,
saved state and outputs are allocated once; ``forward()`` / ``backward()`` only enqueue kernels on the
current stream, so a whole step can be captured into a ``torch.cuda.CUDAGraph`` and replayed.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check
from .functional import _dtype_code, pad_dim, pad_rows

__all__ = ["ContrastiveStep", "PeerStep"]


class ContrastiveStep:
    KERNELS_FORWARD = 3     # prepare, tile kernel (+ zeroes the backward accumulator), finalize (+ column vectors)
    KERNELS_BACKWARD = 2    # tile kernel, finalize

    def __init__(self, loss_kind: int, batch: int, dim: int, temperature: float, normalize: bool = True,
                 dtype: torch.dtype = torch.float32, device="cuda", precision: str = "bf16", deterministic: bool = False):
        self.lib = _lib.load()
        self.flags = _lib.FLAG_DETERMINISTIC if deterministic else 0
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = _lib.PRECISION_SPLIT if precision == "fp32" else _lib.PRECISION_BF16
        self.kind, self.b, self.d = int(loss_kind), int(batch), int(dim)
        self.temperature, self.normalize = float(temperature), bool(normalize)
        self.device = torch.device(device)
        bp, dp = pad_rows(batch), pad_dim(dim)
        dev = self.device
        self.x1 = torch.zeros((batch, dim), dtype=dtype, device=dev)
        self.x2 = torch.zeros((batch, dim), dtype=dtype, device=dev)
        self.code = _dtype_code(self.x1)
        op_bytes = self.lib.simclr_operand_bytes(batch, dim, self.precision)
        if not op_bytes:
            raise ValueError("unsupported shape / precision (fp32-grade operands need d <= 128)")
        self.operand = torch.empty((op_bytes // (2 * dp), dp), dtype=torch.bfloat16, device=dev)
        self.rowvec = torch.empty((4, 2 * bp), dtype=torch.float32, device=dev)   # inv_norm, pos_dot, lse2, row_loss
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.grad1 = torch.empty_like(self.x1)
        self.grad2 = torch.empty_like(self.x2)
        self.fwd_ws_bytes = self.lib.simclr_forward_workspace_bytes(self.kind, batch, batch, dim)
        self.bwd_ws_bytes = self.lib.simclr_backward_workspace_bytes_flags(self.kind, batch, batch, dim, self.flags)
        if not self.fwd_ws_bytes or not self.bwd_ws_bytes:
            raise ValueError("unsupported shape")
        self.fwd_ws = torch.empty(self.fwd_ws_bytes, dtype=torch.uint8, device=dev)
        self.bwd_ws = torch.empty(self.bwd_ws_bytes, dtype=torch.uint8, device=dev)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def prepare(self) -> None:
        """The prepare kernel alone (measurement)."""
        check(self.lib.simclr_prepare_peer(self.kind, self.x1.data_ptr(), self.x2.data_ptr(), self.b, self.d, self.code,
                                           int(self.normalize), self.temperature, self.precision, self.operand.data_ptr(),
                                           self.rowvec[0].data_ptr(), self.rowvec[1].data_ptr(), self.fwd_ws.data_ptr(), 0,
                                           0, None, None, None, self._stream()), "simclr_prepare")

    def forward(self, stage_mask: Optional[int] = None) -> None:
        """prepare + forward.  ``stage_mask`` (measurement only, include/simclr_b200.h SIMCLR_STAGE_*): launch only the
        selected kernels of the forward call, without the prepare kernel, through simclr_forward_stages."""
        lib, st = self.lib, self._stream()
        if stage_mask is None:
            check(lib.simclr_prepare_peer(self.kind, self.x1.data_ptr(), self.x2.data_ptr(), self.b, self.d, self.code,
                                          int(self.normalize), self.temperature, self.precision, self.operand.data_ptr(),
                                          self.rowvec[0].data_ptr(), self.rowvec[1].data_ptr(), self.fwd_ws.data_ptr(), 0, 0,
                                          None, None, None, st), "simclr_prepare")
        # the forward primes the backward workspace (zeroed accumulation buffer, column vectors)
        check(lib.simclr_forward_stages(self.kind, self.operand.data_ptr(), self.operand.data_ptr(), self.b, self.b, 0,
                                        self.d, self.temperature, int(self.normalize), self.precision,
                                        self.rowvec[1].data_ptr(), None, self.rowvec[2].data_ptr(),
                                        self.rowvec[3].data_ptr(), self.stats.data_ptr(), self.loss.data_ptr(),
                                        self.fwd_ws.data_ptr(), self.fwd_ws_bytes, self.bwd_ws.data_ptr(),
                                        self.bwd_ws_bytes, self.x1.data_ptr(), self.x2.data_ptr(), self.code,
                                        self.rowvec[0].data_ptr(), st, _lib.STAGE_ALL if stage_mask is None else stage_mask),
              "simclr_forward")

    def backward(self, grad_out: Optional[torch.Tensor] = None, stage_mask: Optional[int] = None) -> None:
        lib, st = self.lib, self._stream()
        check(lib.simclr_backward_stages(self.kind, self.x1.data_ptr(), self.x2.data_ptr(), self.b, self.b, 0, self.d,
                                         self.code, int(self.normalize), self.temperature, self.precision,
                                         self.operand.data_ptr(), self.operand.data_ptr(), self.rowvec[0].data_ptr(),
                                         self.rowvec[1].data_ptr(), self.rowvec[2].data_ptr(), None,
                                         None if grad_out is None else grad_out.data_ptr(), self.grad1.data_ptr(),
                                         self.grad2.data_ptr(), self.bwd_ws.data_ptr(), self.bwd_ws_bytes,
                                         self.bwd_ws.data_ptr(), self.flags, st,
                                         _lib.STAGE_ALL if stage_mask is None else stage_mask),
              "simclr_backward")

    def step(self, grad_out: Optional[torch.Tensor] = None, x1: Optional[torch.Tensor] = None,
             x2: Optional[torch.Tensor] = None, grad1: Optional[torch.Tensor] = None,
             grad2: Optional[torch.Tensor] = None) -> None:
        """Fused forward+backward (simclr_forward_backward): five launches; the loss statistics are complete when the
        backward finalize kernel has run.  x1 / x2 / grad1 / grad2 default to the runner's own buffers."""
        x1 = self.x1 if x1 is None else x1
        x2 = self.x2 if x2 is None else x2
        grad1 = self.grad1 if grad1 is None else grad1
        grad2 = self.grad2 if grad2 is None else grad2
        check(self.lib.simclr_forward_backward(self.kind, x1.data_ptr(), x2.data_ptr(), self.b, self.d, self.code,
                                               int(self.normalize), self.temperature, self.precision,
                                               None if grad_out is None else grad_out.data_ptr(), self.operand.data_ptr(),
                                               self.rowvec.data_ptr(), self.stats.data_ptr(), self.loss.data_ptr(),
                                               grad1.data_ptr(), grad2.data_ptr(), self.fwd_ws.data_ptr(),
                                               self.fwd_ws_bytes, self.bwd_ws.data_ptr(), self.bwd_ws_bytes, self.flags,
                                               self._stream()),
              "simclr_forward_backward")

    def step_staged(self) -> None:
        self.forward()
        self.backward()


class PeerStep:
    """Allocation-free step of the row-sharded global batch over peer memory (distributed.PeerBatch): what bench.py
    captures into one CUDA graph per rank at N > 1.  No collective call inside: the exchange is NVLink stores from the
    prepare / forward-finalize kernels plus two device-side barriers."""
    KERNELS_FORWARD = 5     # staged: prepare(+push), barrier, tile kernel, finalize(+push), barrier(+statistics)
    KERNELS_BACKWARD = 2    # tile kernel, finalize (the forward primed the workspace)
    KERNELS_FUSED = 5       # step(): prepare(+push), tile kernel(+barrier), finalize(+push), tile kernel(+barrier), finalize

    def __init__(self, loss_kind: int, b_local: int, dim: int, temperature: float, group=None, normalize: bool = True,
                 dtype: torch.dtype = torch.float32, device="cuda"):
        from .distributed import PeerBatch
        self.lib = _lib.load()
        self.kind, self.b, self.d = int(loss_kind), int(b_local), int(dim)
        self.temperature, self.normalize = float(temperature), bool(normalize)
        self.device = torch.device(device)
        self.peer = PeerBatch(b_local, dim, group, self.device)
        bp, dp = pad_rows(b_local), pad_dim(dim)
        dev = self.device
        self.x1 = torch.zeros((b_local, dim), dtype=dtype, device=dev)
        self.x2 = torch.zeros((b_local, dim), dtype=dtype, device=dev)
        self.code = _dtype_code(self.x1)
        self.operand = torch.empty((2 * bp, dp), dtype=torch.bfloat16, device=dev)
        self.rowvec = torch.empty((4, 2 * bp), dtype=torch.float32, device=dev)
        self.stats_local = torch.zeros(4, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.grad1 = torch.empty_like(self.x1)
        self.grad2 = torch.empty_like(self.x2)
        self.fwd_ws_bytes = self.lib.simclr_forward_workspace_bytes(self.kind, b_local, self.peer.b_global, dim)
        self.bwd_ws_bytes = self.lib.simclr_backward_workspace_bytes(self.kind, b_local, self.peer.b_global, dim)
        self.fwd_ws = torch.empty(self.fwd_ws_bytes, dtype=torch.uint8, device=dev)
        self.bwd_ws = torch.empty(self.bwd_ws_bytes, dtype=torch.uint8, device=dev)
        self._cols = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self) -> None:
        self._cols = self.peer.forward(self.kind, self.x1, self.x2, self.temperature, self.normalize, self.operand,
                                       self.rowvec, self.stats_local, self.stats, self.loss, self.fwd_ws,
                                       self.fwd_ws_bytes, self._stream(), self.bwd_ws)

    def backward(self, grad_out: Optional[torch.Tensor] = None) -> None:
        operand_cols, colvec, _gen = self._cols
        p = self.peer
        check(self.lib.simclr_backward(self.kind, self.x1.data_ptr(), self.x2.data_ptr(), self.b, p.b_global, p.row_offset,
                                       self.d, self.code, int(self.normalize), self.temperature, _lib.PRECISION_BF16,
                                       self.operand.data_ptr(),
                                       operand_cols.data_ptr(), self.rowvec[0].data_ptr(), self.rowvec[1].data_ptr(),
                                       None, None, None if grad_out is None else grad_out.data_ptr(),
                                       self.grad1.data_ptr(), self.grad2.data_ptr(), self.bwd_ws.data_ptr(),
                                       self.bwd_ws_bytes, colvec.data_ptr(), 0, self._stream()), "simclr_backward")

    def step(self, grad_out: Optional[torch.Tensor] = None, x1: Optional[torch.Tensor] = None,
             x2: Optional[torch.Tensor] = None, grad1: Optional[torch.Tensor] = None,
             grad2: Optional[torch.Tensor] = None) -> None:
        """Fused step (simclr_forward_backward_peer): five launches, the cross-GPU barriers inside the tile kernels.
        Consecutive steps alternate between the two buffer generations of the PeerBatch: a CUDA graph that is replayed
        must therefore hold an EVEN number of steps (or end with ``barrier()``)."""
        self.peer.fused_step(self.kind, self.x1 if x1 is None else x1, self.x2 if x2 is None else x2, self.temperature,
                             self.normalize, self.operand, self.rowvec, self.stats_local, self.stats, self.loss,
                             self.grad1 if grad1 is None else grad1, self.grad2 if grad2 is None else grad2, self.fwd_ws,
                             self.fwd_ws_bytes, self.bwd_ws, self.bwd_ws_bytes, self._stream(), grad_out)

    def barrier(self) -> None:
        """A stand-alone device-side barrier over the ranks (same flags / epoch as the steps)."""
        p = self.peer
        check(self.lib.simclr_peer_barrier(p.world, p.rank, p._flags, p.epoch.data_ptr(), None, None, None, self._stream()),
              "simclr_peer_barrier")

    def step_staged(self) -> None:
        """prepare / barrier / forward / barrier / backward as separate calls (seven launches)."""
        self.forward()
        self.backward()
