"""RasterEngine -- a lean, allocation-free host driver over the C ABI for the hot loops.

The plugin surface (`diff_gaussian_rasterization`) mirrors the reference one call at a time and therefore
allocates its outputs and opaque buffers per call, like upstream.  The mapping / tracking loops call the same
rasterizer thousands of times with the same shapes, so this engine keeps everything resident: the three opaque
buffers are persistent arenas (the binning arena grows geometrically, never per iteration), image outputs and
the parameter-gradient block are preallocated, and the parameter gradients of all views of an iteration are
summed inside the backward kernel (LVDGS_FLAG_ACCUMULATE) into ONE contiguous float32 block
  [means3D 3 | features 3M | opacity 1 | scales 3 | rotations 4]  (per-array contiguous, back to back)
which is exactly the buffer the multi-GPU mapping step hands to NCCL (no pack kernel).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native
from ._native import RasterParams, ptr

FLAG_EXACT_PP, FLAG_OPACITY_GRAD, FLAG_ACCUMULATE = 1, 2, 4


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


class RasterEngine:
    def __init__(self, P: int, W: int, H: int, sh_coeffs: int = 1, sh_degree: int = 0, device="cuda", flags: int = 0):
        self.L = _native.lib()
        self.dev = torch.device(device)
        self.P, self.W, self.H, self.M, self.D = P, W, H, sh_coeffs, sh_degree
        self.flags = flags
        f32 = dict(dtype=torch.float32, device=self.dev)
        gl, il = _native.GeomLayout(), _native.ImgLayout()
        self.L.lvdgs_get_geom_layout(P, C.byref(gl))
        self.L.lvdgs_get_img_layout(W, H, C.byref(il))
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.arena = {0: torch.empty(gl.total, **u8), 1: torch.empty(1 << 20, **u8), 2: torch.empty(il.total, **u8)}
        self.scratch = torch.empty(self.L.lvdgs_backward_scratch_bytes(P, 0), **u8)
        self._cb = _native.RESIZE_FN(self._resize)
        # outputs
        self.color = torch.empty(3, H, W, **f32)
        self.depth = torch.empty(1, H, W, **f32)
        self.opacity = torch.empty(1, H, W, **f32)
        self.radii = torch.empty(P, dtype=torch.int32, device=self.dev)
        self.n_touched = torch.empty(P, dtype=torch.int32, device=self.dev)
        # gradient block (contiguous; one NCCL message)
        sizes = [("means3D", 3), ("shs", 3 * sh_coeffs), ("opacity", 1), ("scales", 3), ("rotations", 4)]
        total = sum(k for _, k in sizes) * P
        self.grad_flat = torch.zeros(total, **f32)
        self.grads = {}
        off = 0
        for name, k in sizes:
            self.grads[name] = self.grad_flat[off:off + k * P]
            off += k * P
        self.g_means2D = torch.empty(P, 3, **f32)
        self.g_tau = torch.empty(6, **f32)
        self.R = 0
        self.capacity = 0            # instances the binning arena was last laid out for
        self.hint = 0                # speculative-launch capacity hint for the next forward
        if self.dev.index is not None:
            self.L.lvdgs_set_device(self.dev.index)

    def _resize(self, _user, which, nbytes):
        buf = self.arena[int(which)]
        if buf.numel() < nbytes:                       # geometric growth; steady state never allocates
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.dev)
            self.arena[int(which)] = buf
        return buf.data_ptr()

    def _params(self, vc: ViewCamera, flags):
        return RasterParams(P=self.P, sh_degree=self.D, sh_coeffs=self.M, width=vc.W, height=vc.H,
                            tan_fovx=vc.tanfovx, tan_fovy=vc.tanfovy, scale_modifier=1.0, prefiltered=0, debug=0,
                            flags=flags)

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def forward(self, vc: ViewCamera, means3D, opacities, scales, rotations, shs):
        R = C.c_int64(0)
        cap = C.c_int64(0)
        prm = self._params(vc, self.flags)
        rc = self.L.lvdgs_rasterize_forward(C.byref(prm), ptr(vc.bg), ptr(means3D), None, ptr(opacities), ptr(scales),
                                            ptr(rotations), None, ptr(vc.view), ptr(vc.proj), ptr(vc.proj_raw), ptr(shs),
                                            ptr(vc.campos), self._cb, None, C.c_int64(self.hint), ptr(self.color),
                                            ptr(self.radii), ptr(self.depth), ptr(self.opacity), ptr(self.n_touched),
                                            C.byref(R), C.byref(cap), self.stream())
        _native.check(rc, "lvdgs_rasterize_forward")
        self.R = int(R.value)
        self.capacity = int(cap.value)
        self.hint = max(int(self.R * 1.25) + 65536, int(self.hint * 0.98))
        return self.R

    def backward(self, vc: ViewCamera, means3D, opacities, scales, rotations, shs, dL_dcolor, dL_ddepth=None,
                 dL_dopacity=None, accumulate=True):
        flags = self.flags | (FLAG_ACCUMULATE if accumulate else 0)
        prm = self._params(vc, flags)
        g = self.grads
        rc = self.L.lvdgs_rasterize_backward(
            C.byref(prm), ptr(vc.bg), ptr(means3D), ptr(self.radii), None, ptr(opacities), ptr(scales), ptr(rotations),
            None, ptr(vc.view), ptr(vc.proj), ptr(vc.proj_raw), ptr(dL_dcolor), ptr(dL_ddepth), ptr(dL_dopacity), ptr(shs),
            ptr(vc.campos), ptr(self.arena[0]), C.c_int64(self.R), C.c_int64(self.capacity), ptr(self.arena[1]),
            ptr(self.arena[2]),
            ptr(self.scratch), C.c_size_t(self.scratch.numel()), ptr(self.g_means2D), None, ptr(g["opacity"]),
            ptr(g["means3D"]), None, ptr(g["shs"]), ptr(g["scales"]), ptr(g["rotations"]), None,
            ptr(self.g_tau), self.stream())
        _native.check(rc, "lvdgs_rasterize_backward")

    def zero_grads(self):
        self.grad_flat.zero_()

    def pair_count(self):
        """Sum over pixels of n_contrib of the last forward = blended (pixel, Gaussian) pairs the backward visits."""
        il = _native.ImgLayout()
        self.L.lvdgs_get_img_layout(self.W, self.H, C.byref(il))
        nc = self.arena[2][il.n_contrib:il.n_contrib + 4 * self.W * self.H].view(torch.int32)
        return int(nc.sum().item())
