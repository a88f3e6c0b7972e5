"""Read-back of the loss statistics without a stream synchronisation.

The reference API returns the auxiliary-task accuracy as a python float (objective.py:52 / :96), i.e. every call ends
with a device->host read.  ``tensor.item()`` costs a D2H copy plus a stream synchronisation (~45 us on the bench box);
here the finalize kernel writes its four statistics straight into pinned, device-mapped host memory and the caller
polls the last element (written after a system-scope fence) -- the data is there a PCIe write after the kernel ends.

A ring of slots; a slot is reused only after the event recorded behind its previous user has completed.  If polling
does not see the value within a few milliseconds (e.g. a NaN loss, which is indistinguishable from the sentinel), the
event is synchronised instead, which is always correct.
"""
from __future__ import annotations

import math
import threading
import time

import torch

_SLOTS = 64


class _Ring:
    """One ring per device: a CUDA event is bound to the device of its first record, and a slot's event orders the
    slot's reuse against the stream of THAT device."""

    def __init__(self, device: int):
        self.device = device
        self.buf = torch.empty((_SLOTS, 4), dtype=torch.float32).pin_memory()
        self.np = self.buf.numpy()
        self.base = self.buf.data_ptr()
        self.events = [None] * _SLOTS
        self.next = 0
        self.lock = threading.Lock()      # slot hand-out only; a slot is used by the thread that acquired it

    def acquire(self):
        with self.lock:
            i = self.next
            self.next = (i + 1) % _SLOTS
        ev = self.events[i]
        if ev is not None and not ev.query():
            ev.synchronize()              # the previous user of the slot has written it (normally long ago)
        self.np[i, 3] = math.nan          # sentinel: the kernel overwrites it last
        return i

    def pointer(self, i: int) -> int:
        return self.base + 16 * i                    # pinned memory is device-accessible at the same address (UVA)

    def launched(self, i: int) -> None:
        """Record the slot's event behind the launch, on the current stream of the ring's device."""
        ev = self.events[i]
        if ev is None:
            ev = self.events[i] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))

    def abandon(self, i: int) -> None:
        """The call that acquired the slot failed before (or while) launching: nothing will ever write the slot, and a
        stale event of the previous user must not be waited on in its name."""
        self.np[i, 3] = 0.0

    def wait(self, i: int):
        row = self.np[i]
        deadline = None
        spins = 0
        while row[3] != row[3]:           # NaN until the kernel's last store lands
            spins += 1
            if (spins & 255) == 0:
                now = time.perf_counter()
                if deadline is None:
                    deadline = now + 5e-3
                elif now > deadline:
                    if self.events[i] is not None:
                        self.events[i].synchronize()
                    break
        return float(row[0]), float(row[1]), float(row[2]), float(row[3])


_rings = {}
_rings_lock = threading.Lock()


def ring(device=None) -> _Ring:
    """The ring of `device` (index or torch.device; default: the current CUDA device)."""
    if device is None:
        idx = torch.cuda.current_device()
    elif isinstance(device, int):
        idx = device
    else:
        idx = device.index if device.index is not None else torch.cuda.current_device()
    r = _rings.get(idx)
    if r is None:
        with _rings_lock:
            r = _rings.get(idx)
            if r is None:
                r = _rings[idx] = _Ring(idx)
    return r
