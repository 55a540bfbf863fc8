"""Frontend <-> backend map synchronisation without deep copies.

The reference's backend sends the whole Gaussian map to the frontend process after every mapping round:
`clone_obj(self.gaussians)` (utils/multiprocessing_utils.py:21-31: copy.deepcopy + a clone of every tensor attribute)
pushed through an mp.Queue (utils/slam_backend.py:470-480) and adopted by `sync_backend` (utils/slam_frontend.py:1690-
1697) -- 56 B per Gaussian plus Adam state and Python object graphs, re-pickled as fresh CUDA IPC handles every time.

Here the backend owns ONE device allocation of three slots, each holding the rasterizer's inputs of the whole map in
the contiguous block layout of lvdgs.engine.block_layout ([means3D | shs | opacity | scales | rotations], activated
values).  `MapPublisher.publish` copies the current map into a slot no reader holds (one device-to-device pass, the only
data movement), records an inter-process CUDA event and then flips a small header in host shared memory (version, latest
slot, number of Gaussians per slot).  The frontend opens the allocation ONCE (`MapSubscriber(handle)`, the handle
travels through the existing mp.Queue) and `acquire()` returns views into the latest slot -- no copy, no pickling -- after
making its stream wait on the publisher's event; `release()` hands the slot back.  Three slots mean neither side ever
waits for the other: one slot may be held by the reader, one is the latest published, one is free for the next publish.
The views quack like the GaussianModel attributes `gaussian_renderer.render` reads (get_xyz, get_opacity, ...).

Works on CPU tensors too (host shared memory, no events): that is what the `-m "not gpu"` tests exercise.
"""
from __future__ import annotations

import fcntl
import os
import tempfile
from typing import Dict, Optional

import torch

from .engine import GROUPS, block_layout

N_SLOTS = 3
_LOCAL_EVENTS = {}      # lock path -> the publisher's events, for a subscriber living in the publisher's own process (CUDA
                        # refuses to open an IPC event handle in the process that exported it)
# header (int64, host shared memory): [version, latest slot, reader slot (-1: none), P of slot 0..2, capacity, sh_coeffs]
_VERSION, _LATEST, _READER, _P0, _CAPACITY, _M = 0, 1, 2, 3, 6, 7


class _HeaderLock:
    """Advisory lock on a file both processes can name (no inheritance constraints, unlike mp.Lock through a Queue)."""

    def __init__(self, path):
        self.path = path
        self.f = None

    def __enter__(self):
        self.f = open(self.path, "a+")
        fcntl.flock(self.f, fcntl.LOCK_EX)

    def __exit__(self, *exc):
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


class MapView:
    """What render() reads from a GaussianModel, as views into one published slot."""

    def __init__(self, flat: torch.Tensor, P: int, M: int, version: int, slot: int, active_sh_degree: int = 0):
        layout, _ = block_layout(P, M)
        v = lambda name, *shape: flat[layout[name][0]:layout[name][0] + layout[name][1]].view(*shape)
        self.get_xyz = v("means3D", P, 3)
        self.get_features = v("shs", P, M, 3)
        self.get_opacity = v("opacity", P, 1)
        self.get_scaling = v("scales", P, 3)
        self.get_rotation = v("rotations", P, 4)
        self.active_sh_degree = active_sh_degree
        self.P, self.version, self.slot = P, version, slot


class MapPublisher:
    def __init__(self, capacity: int, sh_coeffs: int = 1, device="cuda", active_sh_degree: int = 0):
        self.device = torch.device(device)
        self.capacity, self.M, self.active_sh_degree = int(capacity), int(sh_coeffs), int(active_sh_degree)
        _, self.slot_len = block_layout(self.capacity, self.M)
        self.buf = torch.zeros(N_SLOTS, self.slot_len, dtype=torch.float32, device=self.device)
        if not self.buf.is_cuda:
            self.buf.share_memory_()
        self.header = torch.zeros(8, dtype=torch.int64).share_memory_()
        self.header[_LATEST], self.header[_READER], self.header[_CAPACITY], self.header[_M] = -1, -1, self.capacity, self.M
        fd, self.lock_path = tempfile.mkstemp(prefix="lvdgs_mapsync_", suffix=".lock")
        os.close(fd)
        self.lock = _HeaderLock(self.lock_path)
        self.events = [torch.cuda.Event(interprocess=True) for _ in range(N_SLOTS)] if self.buf.is_cuda else None
        if self.events is not None:
            _LOCAL_EVENTS[self.lock_path] = (os.getpid(), self.events)

    def handle(self) -> Dict:
        """Picklable description for the subscriber process (send it through the frontend queue once)."""
        h = dict(buf=self.buf, header=self.header, lock_path=self.lock_path, M=self.M, capacity=self.capacity,
                 active_sh_degree=self.active_sh_degree, device=str(self.device))
        if self.events is not None:
            h["events"] = [e.ipc_handle() for e in self.events]
        return h

    def publish(self, arrays: Dict[str, torch.Tensor]) -> int:
        """arrays: the five rasterizer inputs (means3D [P,3], shs [P,M,3], opacity [P,1] or [P], scales [P,3], rotations
        [P,4]; e.g. {n: mapper.view(n) for n in GROUPS}).  One device-to-device pass; returns the new version."""
        P = int(arrays["means3D"].shape[0])
        if P > self.capacity:
            raise ValueError(f"MapPublisher: {P} Gaussians exceed the capacity {self.capacity} the allocation was opened with")
        with self.lock:
            busy = {int(self.header[_LATEST]), int(self.header[_READER])}
        slot = next(s for s in range(N_SLOTS) if s not in busy)
        layout, _ = block_layout(P, self.M)
        dst = self.buf[slot]
        for name in GROUPS:
            off, ln = layout[name]
            dst[off:off + ln].copy_(arrays[name].detach().reshape(-1), non_blocking=True)
        if self.events is not None:
            self.events[slot].record(torch.cuda.current_stream(self.device))
        with self.lock:
            self.header[_P0 + slot] = P
            self.header[_LATEST] = slot
            self.header[_VERSION] += 1
            return int(self.header[_VERSION])

    def close(self):
        _LOCAL_EVENTS.pop(self.lock_path, None)
        try:
            os.unlink(self.lock_path)
        except OSError:
            pass


class MapSubscriber:
    def __init__(self, handle: Dict):
        self.buf, self.header = handle["buf"], handle["header"]
        self.M, self.active_sh_degree = int(handle["M"]), int(handle["active_sh_degree"])
        self.lock = _HeaderLock(handle["lock_path"])
        self.device = self.buf.device
        self.events = None
        if "events" in handle:
            local = _LOCAL_EVENTS.get(handle["lock_path"])
            if local is not None and local[0] == os.getpid():
                self.events = local[1]
            else:
                self.events = [torch.cuda.Event.from_ipc_handle(self.device, h) for h in handle["events"]]
        self.held: Optional[int] = None

    def version(self) -> int:
        return int(self.header[_VERSION])

    def acquire(self) -> Optional[MapView]:
        """Views into the latest published map (None before the first publish).  Holds the slot until release()."""
        if self.held is not None:
            self.release()
        with self.lock:
            slot = int(self.header[_LATEST])
            if slot < 0:
                return None
            self.header[_READER] = slot
            P, version = int(self.header[_P0 + slot]), int(self.header[_VERSION])
        self.held = slot
        if self.events is not None:
            self.events[slot].wait(torch.cuda.current_stream(self.device))      # the publisher's copy precedes our reads
        return MapView(self.buf[slot], P, self.M, version, slot, self.active_sh_degree)

    def release(self):
        """Call when the renders that use the acquired views have been enqueued; waits for them, then frees the slot."""
        if self.held is None:
            return
        if self.buf.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()
        with self.lock:
            self.header[_READER] = -1
        self.held = None
