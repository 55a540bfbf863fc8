"""RasterEngine -- a lean, allocation-free host driver over the C ABI for the hot loops.

The plugin surface (`diff_gaussian_rasterization`) mirrors the reference one call at a time and therefore
allocates its outputs and opaque buffers per call, like upstream.  The mapping / tracking loops call the same
rasterizer thousands of times with the same shapes, so this engine keeps everything resident: the three opaque
buffers are persistent arenas (the binning arena grows geometrically, never per iteration), image outputs and
the parameter-gradient block are preallocated, and the parameter gradients of all views of an iteration are
summed inside the backward kernel (LVDGS_FLAG_ACCUMULATE) into ONE contiguous float32 block
  [means3D 3 | features 3M | opacity 1 | scales 3 | rotations 4]  (per-array contiguous, back to back)
which is exactly the buffer the multi-GPU mapping step hands to NCCL (no pack kernel).

Views of one mapping iteration are independent given the map (utils/slam_backend.py:180-300 renders them one after
the other only because the reference has one stream).  `run_views` therefore software-pipelines them over two CUDA
streams and two buffer slots: the forward of view k+1 (preprocess, binning, six latency-bound sort passes, blend)
runs on the forward stream while the backward of view k (issue-bound blend backward) runs on the backward stream;
all backwards stay on one stream, so the accumulation into the gradient block is race-free.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native
from ._native import RasterParams, ptr

FLAG_EXACT_PP, FLAG_OPACITY_GRAD, FLAG_ACCUMULATE, FLAG_POSE_ONLY = 1, 2, 4, 8
FLAG_ZEROED_SCRATCH = 64

# parameter groups of the contiguous parameter / gradient block, in block order
GROUPS = ("means3D", "shs", "opacity", "scales", "rotations")


def block_layout(P: int, sh_coeffs: int = 1, multiple: int = 4):
    """Offsets (in floats) of the five per-Gaussian arrays inside ONE contiguous float32 block -- the layout shared by
    RasterEngine.grad_flat and lvdgs.mapping.ShardedMapper.param_flat: [means3D 3 | shs 3M | opacity 1 | scales 3 |
    rotations 4] per Gaussian, array after array, every array starting on a 16-byte boundary (float4 accesses in the
    kernels) and the total rounded up to `multiple` floats (a multiple of 4 x world makes the block reduce-scatterable).
    Returns ({name: (offset, length)}, total)."""
    widths = {"means3D": 3, "shs": 3 * sh_coeffs, "opacity": 1, "scales": 3, "rotations": 4}
    out, off = {}, 0
    for name in GROUPS:
        n = widths[name] * P
        out[name] = (off, n)
        off += (n + 3) & ~3
    multiple = max(4, int(multiple))
    total = (off + multiple - 1) // multiple * multiple
    return out, total


class ViewCamera:
    """Device-resident camera block: the five small tensors GaussianRasterizationSettings carries."""

    def __init__(self, cam, device, bg=(0.0, 0.0, 0.0)):
        t = lambda a: torch.tensor(a, dtype=torch.float32, device=device).contiguous()
        self.W, self.H = int(cam.image_width), int(cam.image_height)
        self.tanfovx, self.tanfovy = float(cam.tanfovx), float(cam.tanfovy)
        self.view = t(cam.world_view_transform)
        self.proj = t(cam.full_proj_transform)
        self.proj_raw = t(cam.projection_matrix)
        self.campos = t(cam.camera_center)
        self.bg = t(bg)


class _Slot:
    """One in-flight render: opaque arenas, image outputs, per-view gradient outputs."""

    def __init__(self, eng):
        L, P, W, H, dev = eng.L, eng.P, eng.W, eng.H, eng.dev
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        gl, il = _native.GeomLayout(), _native.ImgLayout()
        L.lvdgs_get_geom_layout(P, C.byref(gl))
        L.lvdgs_get_img_layout(W, H, C.byref(il))
        self.dev = dev
        self.arena = {0: torch.empty(gl.total, **u8), 1: torch.empty(1 << 20, **u8), 2: torch.empty(il.total, **u8)}
        self.scratch = torch.empty(L.lvdgs_backward_scratch_bytes(P, 0), **u8)
        self.cb = _native.RESIZE_FN(self._resize)
        self.color = torch.empty(3, H, W, **f32)
        self.depth = torch.empty(1, H, W, **f32)
        self.opacity = torch.empty(1, H, W, **f32)
        self._cap_P = P
        self._radii = torch.empty(P, dtype=torch.int32, device=dev)
        self._n_touched = torch.empty(P, dtype=torch.int32, device=dev)
        self._g_means2D = torch.empty(P, 3, **f32)
        self.radii, self.n_touched, self.g_means2D = self._radii, self._n_touched, self._g_means2D
        self.g_tau = torch.empty(6, **f32)
        self.streams = ()            # side streams that use this slot's tensors (set by the engine)
        self.R = 0
        self.scratch_clean = False   # the forward cleared the backward's accumulator rows (LVDGS_FLAG_ZEROED_SCRATCH)
        self.capacity = 0            # instances the binning arena was last laid out for
        self.hint = 0                # speculative-launch capacity hint for the next forward

    def set_num_gaussians(self, L, P: int):
        """Per-Gaussian arrays of the slot for a map of P Gaussians: views into allocations that only ever grow (1.25x), so
        densify / prune churn between iterations does not allocate in the steady state."""
        f32 = dict(dtype=torch.float32, device=self.dev)
        if P > self._cap_P:
            for t in (self._radii, self._n_touched, self._g_means2D, self.scratch):
                for st in self.streams:
                    t.record_stream(st)
            self._cap_P = int(P * 1.25) + 1024
            self._radii = torch.empty(self._cap_P, dtype=torch.int32, device=self.dev)
            self._n_touched = torch.empty(self._cap_P, dtype=torch.int32, device=self.dev)
            self._g_means2D = torch.empty(self._cap_P, 3, **f32)
            self.scratch = torch.empty(L.lvdgs_backward_scratch_bytes(self._cap_P, 0), dtype=torch.uint8, device=self.dev)
            for t in (self._radii, self._n_touched, self._g_means2D, self.scratch):
                for st in self.streams:
                    t.record_stream(st)
        self.radii, self.n_touched, self.g_means2D = self._radii[:P], self._n_touched[:P], self._g_means2D[:P]

    def _resize(self, _user, which, nbytes):
        buf = self.arena[int(which)]
        if buf.numel() < nbytes:                       # geometric growth; steady state never allocates
            # the replaced arena may still be read by this slot's previous backward on the backward stream (the host
            # never blocks on it): tell the caching allocator, so the block is not handed out again before those kernels
            # have finished
            for st in self.streams:
                buf.record_stream(st)
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.dev)
            for st in self.streams:
                buf.record_stream(st)
            self.arena[int(which)] = buf
        return buf.data_ptr()


class RasterEngine:
    def __init__(self, P: int, W: int, H: int, sh_coeffs: int = 1, sh_degree: int = 0, device="cuda", flags: int = 0,
                 slots: int = 2, grad_flat=None):
        self.L = _native.lib()
        self.dev = torch.device(device)
        self.P, self.W, self.H, self.M, self.D = P, W, H, sh_coeffs, sh_degree
        self.flags = flags
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.slots = [_Slot(self) for _ in range(max(1, slots))]
        # gradient block (contiguous; one NCCL message).  A caller that shards the optimiser (lvdgs.mapping) passes its own,
        # suitably padded, buffer.
        layout, total = block_layout(P, sh_coeffs)
        if grad_flat is not None:
            assert grad_flat.numel() >= total and grad_flat.dtype == torch.float32 and grad_flat.is_contiguous()
        self.grad_flat = grad_flat if grad_flat is not None else torch.zeros(total, **f32)
        self.grads = {name: self.grad_flat[off:off + n] for name, (off, n) in layout.items()}
        import os
        self.s_fwd = torch.cuda.Stream(self.dev, priority=int(os.environ.get("LVDGS_FWD_PRIORITY", "-1")))   # short latency-bound kernels get SM slots first
        self.s_bwd = torch.cuda.Stream(self.dev)
        # every slot tensor is allocated on the construction stream but used on the two side streams
        for sl in self.slots:
            sl.streams = (self.s_fwd, self.s_bwd)
            for t in (*sl.arena.values(), sl.scratch, sl.color, sl.depth, sl.opacity, sl.radii, sl.n_touched, sl.g_means2D, sl.g_tau):
                for st in sl.streams:
                    t.record_stream(st)
        if self.dev.index is not None:
            self.L.lvdgs_set_device(self.dev.index)

    def set_num_gaussians(self, P: int, grad_flat=None):
        """The map changed size (densify / prune): re-point the per-Gaussian arrays and the gradient block.  `grad_flat`: the
        caller's block for the new size (lvdgs.mapping.ShardedMapper.new_grad_block()); None allocates one here."""
        self.P = int(P)
        for sl in self.slots:
            sl.set_num_gaussians(self.L, self.P)
        layout, total = block_layout(self.P, self.M)
        if grad_flat is None:
            grad_flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        assert grad_flat.numel() >= total
        self.grad_flat = grad_flat
        self.grads = {name: self.grad_flat[off:off + n] for name, (off, n) in layout.items()}

    # ---- slot 0 shortcuts (single-view use) ----
    color = property(lambda self: self.slots[0].color)
    depth = property(lambda self: self.slots[0].depth)
    opacity = property(lambda self: self.slots[0].opacity)
    radii = property(lambda self: self.slots[0].radii)
    n_touched = property(lambda self: self.slots[0].n_touched)
    g_means2D = property(lambda self: self.slots[0].g_means2D)
    g_tau = property(lambda self: self.slots[0].g_tau)
    R = property(lambda self: self.slots[0].R)
    arena = property(lambda self: self.slots[0].arena)

    def _params(self, vc, flags):
        return RasterParams(P=self.P, sh_degree=self.D, sh_coeffs=self.M, width=vc.W, height=vc.H,
                            tan_fovx=vc.tanfovx, tan_fovy=vc.tanfovy, scale_modifier=1.0, prefiltered=0, debug=0,
                            flags=flags)

    def _stream(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        return C.c_void_p(s.cuda_stream)

    def forward(self, vc, means3D, opacities, scales, rotations, shs, slot: int = 0, stream=None):
        sl = self.slots[slot]
        R = C.c_int64(0)
        cap = C.c_int64(0)
        prm = self._params(vc, self.flags)
        # the backward's accumulator rows are cleared on the FORWARD's stream, ahead of its kernels (LVDGS_FLAG_ZEROED_SCRATCH):
        # in run_views that is off the backward stream, which carries the critical path, and on a single stream (tracking) it
        # keeps the kernel chain forward -> loss -> backward free of memsets (programmatic dependent launches)
        self.L.lvdgs_zero_async(ptr(sl.scratch), C.c_size_t(self.L.lvdgs_backward_scratch_bytes(self.P, 0)), self._stream(stream))
        rc = self.L.lvdgs_rasterize_forward(C.byref(prm), ptr(vc.bg), ptr(means3D), None, ptr(opacities), ptr(scales),
                                            ptr(rotations), None, ptr(vc.view), ptr(vc.proj), ptr(vc.proj_raw), ptr(shs),
                                            ptr(vc.campos), sl.cb, None, C.c_int64(sl.hint), ptr(sl.color),
                                            ptr(sl.radii), ptr(sl.depth), ptr(sl.opacity), ptr(sl.n_touched),
                                            C.byref(R), C.byref(cap), self._stream(stream))
        _native.check(rc, "lvdgs_rasterize_forward")
        sl.scratch_clean = True
        sl.R = int(R.value)
        sl.capacity = int(cap.value)
        sl.hint = max(int(sl.R * 1.25) + 65536, int(sl.hint * 0.98))
        return sl.R

    def backward(self, vc, means3D, opacities, scales, rotations, shs, dL_dcolor, dL_ddepth=None,
                 dL_dopacity=None, accumulate=True, slot: int = 0, stream=None, pose_only: bool = False):
        """pose_only (tracking): only slot.g_tau is produced, no parameter gradients and no screen-space gradients."""
        sl = self.slots[slot]
        flags = self.flags | (FLAG_ACCUMULATE if accumulate and not pose_only else 0) | (FLAG_POSE_ONLY if pose_only else 0)
        if sl.scratch_clean:            # one backward per forward consumes the cleared scratch
            flags |= FLAG_ZEROED_SCRATCH
            sl.scratch_clean = False
        prm = self._params(vc, flags)
        g = {k: None for k in self.grads} if pose_only else self.grads
        rc = self.L.lvdgs_rasterize_backward(
            C.byref(prm), ptr(vc.bg), ptr(means3D), ptr(sl.radii), None, ptr(opacities), ptr(scales), ptr(rotations),
            None, ptr(vc.view), ptr(vc.proj), ptr(vc.proj_raw), ptr(dL_dcolor), ptr(dL_ddepth), ptr(dL_dopacity), ptr(shs),
            ptr(vc.campos), ptr(sl.arena[0]), C.c_int64(sl.R), C.c_int64(sl.capacity), ptr(sl.arena[1]),
            ptr(sl.arena[2]), ptr(sl.scratch), C.c_size_t(sl.scratch.numel()), None if pose_only else ptr(sl.g_means2D), None, ptr(g["opacity"]),
            ptr(g["means3D"]), None, ptr(g["shs"]), ptr(g["scales"]), ptr(g["rotations"]), None,
            ptr(sl.g_tau), self._stream(stream))
        _native.check(rc, "lvdgs_rasterize_backward")

    def view_streams(self, n_views: int):
        """(forward stream, backward stream) run_views uses for a window of n_views: the two side streams, or the caller's
        current stream for a single view (callbacks that queue waits on "the forward stream" ask here)."""
        if n_views == 1:
            cur = torch.cuda.current_stream(self.dev)
            return cur, cur
        return self.s_fwd, self.s_bwd

    def run_views(self, vcs, means3D, opacities, scales, rotations, shs, upstream, on_view=None, before_view=None, bwd_wait=None):
        """Forward + backward of several views with gradients accumulated into `grad_flat`, software-pipelined over the
        forward / backward streams and the buffer slots.  `upstream(k, slot)` is called on the backward stream after
        view k's forward has finished and returns (dL_dcolor, dL_ddepth, dL_dopacity) for it -- the place where a caller
        computes its loss from `slot.color/depth/opacity`.  `on_view(k, slot)` (optional) runs on the backward stream
        after view k's backward (e.g. densification statistics from slot.g_means2D / slot.radii).  `before_view(k)`
        (optional) runs on the host right before view k's forward is queued.  `bwd_wait` (optional): an event the first
        backward has to wait for (the forwards do not)."""
        cur = torch.cuda.current_stream(self.dev)
        if len(vcs) == 1:
            # one view (a rank of an 8-GPU window): nothing to overlap, so no stream hops -- forward, loss and backward run
            # back to back on the caller's stream (three cross-stream dependencies less on the critical path)
            if bwd_wait is not None:
                cur.wait_event(bwd_wait)
            if before_view is not None:
                before_view(0)
            self.forward(vcs[0], means3D, opacities, scales, rotations, shs, slot=0, stream=cur)
            gc, gd, go = upstream(0, self.slots[0])
            self.backward(vcs[0], means3D, opacities, scales, rotations, shs, gc, gd, go, accumulate=True, slot=0, stream=cur)
            if on_view is not None:
                on_view(0, self.slots[0])
            return
        n = len(self.slots)
        start = torch.cuda.Event(); start.record(cur)
        self.s_fwd.wait_event(start); self.s_bwd.wait_event(start)
        if bwd_wait is not None:       # e.g. the mapper's deferred clearing of the gradient block (ShardedMapper.grad_ready)
            self.s_bwd.wait_event(bwd_wait)
        bwd_done = [None] * n
        for k, vc in enumerate(vcs):
            s = k % n
            if bwd_done[s] is not None:
                self.s_fwd.wait_event(bwd_done[s])                 # slot s is free again
            if before_view is not None:
                before_view(k)                                     # e.g. make the forward stream wait for view k's camera upload
            self.forward(vc, means3D, opacities, scales, rotations, shs, slot=s, stream=self.s_fwd)
            fwd_done = torch.cuda.Event(); fwd_done.record(self.s_fwd)
            self.s_bwd.wait_event(fwd_done)
            with torch.cuda.stream(self.s_bwd):
                gc, gd, go = upstream(k, self.slots[s])
                self.backward(vc, means3D, opacities, scales, rotations, shs, gc, gd, go, accumulate=True, slot=s,
                              stream=self.s_bwd)
                if on_view is not None:
                    on_view(k, self.slots[s])
            bwd_done[s] = torch.cuda.Event(); bwd_done[s].record(self.s_bwd)
        end = torch.cuda.Event(); end.record(self.s_bwd)
        cur.wait_event(end)
        endf = torch.cuda.Event(); endf.record(self.s_fwd)
        cur.wait_event(endf)

    def zero_grads(self):
        self.grad_flat.zero_()

    def pair_count(self, slot: int = 0):
        """Sum over pixels of n_contrib of the last forward = blended (pixel, Gaussian) pairs the backward visits."""
        il = _native.ImgLayout()
        self.L.lvdgs_get_img_layout(self.W, self.H, C.byref(il))
        nc = self.slots[slot].arena[2][il.n_contrib:il.n_contrib + 4 * self.W * self.H].view(torch.int32)
        return int(nc.sum().item())
