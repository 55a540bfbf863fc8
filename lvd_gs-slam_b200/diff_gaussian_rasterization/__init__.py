"""Drop-in replacement for the reference's `diff_gaussian_rasterization` plugin (MonoGS "-w-pose" fork).

Same Python surface as the module `pip install submodules/diff-gaussian-rasterization` provides
(/root/reference/README.md:39-44): `GaussianRasterizationSettings`, `GaussianRasterizer` and the autograd
function behind it, imported by the reference's `gaussian_splatting.gaussian_renderer.render`
(call sites utils/slam_frontend.py:1493, utils/slam_backend.py:98,184,277,407, utils/eval_utils_0806.py:215,
utils/init_pose.py:145).  Underneath it is a thin ctypes binding over the `extern "C"` launchers of
liblvdgs.so (include/lvdgs.h): hand-written sm_100a CUDA, no CPU fallback -- a missing library raises.

Returned tuple: (color [3,H,W], radii [P] int32, depth [1,H,W], opacity [1,H,W], n_touched [P] int32).
Gradients: means3D, means2D (screen-space, xy used), sh / colors_precomp, opacities, scales, rotations,
cov3D_precomp, theta, rho (camera pose, tau = [rho; theta] of utils/pose_utils.py:70-87).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import NamedTuple

import torch
import torch.nn as nn

from lvdgs import _native

# Upstream drops grad_out_opacity and omits the principal point in the pose Jacobian.
# Set LVDGS_FLAGS=1|2 (or assign module attribute FLAGS) to use the true derivatives instead.
FLAGS = int(os.environ.get("LVDGS_FLAGS", "0"))

# Speculative-launch capacity hints (include/lvdgs.h: lvdgs_rasterize_forward): per device, the next forward sizes its
# binning buffer for 1.25x the largest instance count seen recently, so the device never idles while the host reads R.
# LVDGS_SPECULATIVE=0 restores upstream's read-R-then-launch order.
SPECULATIVE = os.environ.get("LVDGS_SPECULATIVE", "1") != "0"
_capacity_hint = {}


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _prep(t, device=None):
    """float32, contiguous, 16-byte aligned CUDA tensor or None (None / empty tensors mean 'not provided')."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _make_resizer(device, store):
    """ctypes resize callback for the three opaque buffers of the C ABI.  `store` (a plain dict: which -> uint8 tensor)
    is what the autograd ctx keeps alive; the callback object itself is dropped right after the forward call, so no
    reference cycle delays the release of the buffers (they are tens of MB each)."""

    def _resize(_user, which, nbytes):
        buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
        store[int(which)] = buf
        return buf.data_ptr()

    return _native.RESIZE_FN(_resize)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta,
                        rho, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, theta, rho, raster_settings)


def _params(rs: GaussianRasterizationSettings, P: int, M: int) -> _native.RasterParams:
    return _native.RasterParams(P=P, sh_degree=int(rs.sh_degree), sh_coeffs=M, width=int(rs.image_width),
                                height=int(rs.image_height), tan_fovx=float(rs.tanfovx), tan_fovy=float(rs.tanfovy),
                                scale_modifier=float(rs.scale_modifier), prefiltered=int(bool(rs.prefiltered)),
                                debug=int(bool(rs.debug)), flags=FLAGS)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                raster_settings):
        rs = raster_settings
        if not means3D.is_cuda:
            raise RuntimeError("diff_gaussian_rasterization (B200): tensors must be on a CUDA device; there is no CPU path")
        L = _native.lib()
        dev = means3D.device
        m3 = _prep(means3D)
        P = 0 if m3 is None else m3.shape[0]
        shs = _prep(sh); cp = _prep(colors_precomp); op = _prep(opacities)
        sc = _prep(scales); rot = _prep(rotations); cov = _prep(cov3Ds_precomp)
        M = 0 if shs is None else shs.shape[1]
        H, W = int(rs.image_height), int(rs.image_width)
        bg = _prep(rs.bg.to(dev)); view = _prep(rs.viewmatrix.to(dev)); proj = _prep(rs.projmatrix.to(dev))
        praw = _prep(rs.projmatrix_raw.to(dev)); campos = _prep(rs.campos.to(dev))
        color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
        depth = torch.empty(1, H, W, dtype=torch.float32, device=dev)
        opac_img = torch.empty(1, H, W, dtype=torch.float32, device=dev)
        radii = torch.empty(P, dtype=torch.int32, device=dev)
        n_touched = torch.empty(P, dtype=torch.int32, device=dev)
        bufs = {}
        resize_cb = _make_resizer(dev, bufs)
        prm = _params(rs, P, M)
        R = C.c_int64(0)
        cap = C.c_int64(0)
        hint = _capacity_hint.get(dev.index, 0) if SPECULATIVE else 0
        if dev.index is not None:
            L.lvdgs_set_device(dev.index)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = _native.ptr
        rc = L.lvdgs_rasterize_forward(C.byref(prm), p(bg), p(m3), p(cp), p(op), p(sc), p(rot), p(cov), p(view), p(proj),
                                       p(praw), p(shs), p(campos), resize_cb, None, C.c_int64(hint), p(color), p(radii),
                                       p(depth), p(opac_img), p(n_touched), C.byref(R), C.byref(cap), stream)
        del resize_cb
        _native.check(rc, "lvdgs_rasterize_forward")
        ctx.rs = rs
        ctx.num_rendered = int(R.value)
        ctx.capacity = int(cap.value)
        if SPECULATIVE:   # decay slowly, grow at once
            _capacity_hint[dev.index] = max(int(R.value * 1.25) + 65536, int(hint * 0.98))
        ctx.bufs = bufs
        ctx.shapes = (P, M)
        ctx.aux = (bg, view, proj, praw, campos)
        ctx.in_shapes = (None if theta is None else theta.shape, None if rho is None else rho.shape)
        ctx.save_for_backward(m3, shs, cp, op, sc, rot, cov, radii)
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, opac_img, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_out_depth, grad_out_opacity, grad_n_touched):
        L = _native.lib()
        rs = ctx.rs
        P, M = ctx.shapes
        m3, shs, cp, op, sc, rot, cov, radii = ctx.saved_tensors
        bg, view, proj, praw, campos = ctx.aux
        dev = grad_out_color.device
        f32 = dict(dtype=torch.float32, device=dev)
        g_means2D = torch.empty(P, 3, **f32); g_opac = torch.empty(P, 1, **f32); g_means3D = torch.empty(P, 3, **f32)
        g_colors = torch.empty(P, 3, **f32) if cp is not None else None
        g_cov = torch.empty(P, 6, **f32) if cov is not None else None
        g_sh = torch.empty(P, M, 3, **f32) if shs is not None else None
        g_sc = torch.empty(P, 3, **f32) if cov is None else None
        g_rot = torch.empty(P, 4, **f32) if cov is None else None
        g_tau = torch.empty(6, **f32)
        scratch = torch.empty(L.lvdgs_backward_scratch_bytes(P, ctx.num_rendered), dtype=torch.uint8, device=dev)
        prm = _params(rs, P, M)
        if dev.index is not None:
            L.lvdgs_set_device(dev.index)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = _native.ptr
        gc = _prep(grad_out_color)
        gd = _prep(grad_out_depth) if grad_out_depth is not None else None
        go = _prep(grad_out_opacity) if (grad_out_opacity is not None and (FLAGS & 2)) else None
        b = ctx.bufs
        rc = L.lvdgs_rasterize_backward(C.byref(prm), p(bg), p(m3), p(radii), p(cp), p(op), p(sc), p(rot), p(cov), p(view),
                                        p(proj), p(praw), p(gc), p(gd), p(go), p(shs), p(campos), p(b.get(0)),
                                        C.c_int64(ctx.num_rendered), C.c_int64(ctx.capacity), p(b.get(1)), p(b.get(2)),
                                        p(scratch),
                                        C.c_size_t(scratch.numel()), p(g_means2D), p(g_colors), p(g_opac), p(g_means3D),
                                        p(g_cov), p(g_sh), p(g_sc), p(g_rot), None, p(g_tau), stream)
        _native.check(rc, "lvdgs_rasterize_backward")
        th_shape, rho_shape = ctx.in_shapes
        g_rho = g_tau[:3].reshape(rho_shape) if rho_shape is not None else None
        g_theta = g_tau[3:].reshape(th_shape) if th_shape is not None else None
        return (g_means3D, g_means2D, g_sh, g_colors, g_opac, g_sc, g_rot, g_cov, g_theta, g_rho, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            L = _native.lib()
            pos = _prep(positions)
            P = positions.shape[0]
            out = torch.empty(P, dtype=torch.uint8, device=positions.device)
            view = _prep(rs.viewmatrix.to(positions.device)); proj = _prep(rs.projmatrix.to(positions.device))
            stream = C.c_void_p(torch.cuda.current_stream(positions.device).cuda_stream)
            rc = L.lvdgs_mark_visible(P, _native.ptr(pos), _native.ptr(view), _native.ptr(proj), _native.ptr(out), stream)
            _native.check(rc, "lvdgs_mark_visible")
        return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        theta = empty if theta is None else theta
        rho = empty if rho is None else rho
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   theta, rho, rs)
