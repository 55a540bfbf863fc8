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
import threading
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
# LVDGS_ZEROED_OUTPUTS=1: the backward zero-fills its gradient block and lets the library skip the culled Gaussians
# (LVDGS_FLAG_ZEROED_OUTPUTS).  Off by default: measured on B200 at 500k Gaussians the 28 MB fill costs more than the
# shorter kernel saves (0.637-0.657 vs 0.614 ms per fwd+bwd); the flag pays when the caller's buffers are zero anyway.
ZEROED_OUTPUTS = os.environ.get("LVDGS_ZEROED_OUTPUTS", "0") != "0"


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


def _prep(t):
    """float32, contiguous, 16-byte aligned CUDA tensor or None (None / empty tensors mean 'not provided')."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype is torch.float32 and t.is_contiguous() and not (t.data_ptr() & 15):
        return t
    t = t.contiguous().float()
    return t.clone() if t.data_ptr() & 15 else t


def _on(t, dev):
    return t if t.device == dev else t.to(dev)


def _ptr(t):
    return None if t is None else t.data_ptr()


# One persistent ctypes callback per process; the per-call destination (the dict the autograd ctx keeps alive) is
# thread-local, so no callback object -- and no reference cycle through it -- is created per render.
_tls = threading.local()


def _resize(_user, which, nbytes):
    buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=_tls.dev)
    _tls.store[int(which)] = buf
    return buf.data_ptr()


_RESIZE_CB = _native.RESIZE_FN(_resize)

# Host-side cost of a render matters: the reference calls the plugin 100 times per tracked frame.  The three opaque
# buffers are therefore carved out of ONE allocation whose sizes are known before the call (lvdgs_get_*_layout, cached
# per shape) and handed to the library through its own C callback (lvdgs_static_resize) -- no call back into Python per
# buffer; only a request that does not fit (a repeated speculative tail) reaches `_resize` above.
_layout_cache = {}            # (P, W, H) -> (geom bytes, img bytes);  capacity -> binning bytes
_static_cb = None


def debug_buffers(node):
    """{LVDGS_BUF_GEOM / _BINNING / _IMG: uint8 tensor} of a forward, from its autograd node (`color.grad_fn`): the three
    opaque buffers as lvdgs._native.debug_views reads them (parity tests / debugging only)."""
    arena, geom, binning, img, extra = node.arena
    out = {}
    for which, (off, n) in ((0, geom), (1, binning), (2, img)):
        if which in extra:
            out[which] = extra[which]
        elif n:
            out[which] = arena[off:off + n]
    return out


def _layout_sizes(L, P, W, H):
    key = (P, W, H)
    v = _layout_cache.get(key)
    if v is None:
        gl, il = _native.GeomLayout(), _native.ImgLayout()
        L.lvdgs_get_geom_layout(P, C.byref(gl)); L.lvdgs_get_img_layout(W, H, C.byref(il))
        if len(_layout_cache) > 256:
            _layout_cache.clear()
        v = _layout_cache[key] = (int(gl.total), int(il.total))
    return v


def _binning_size(L, capacity):
    v = _layout_cache.get(capacity)
    if v is None:
        bl = _native.BinningLayout()
        L.lvdgs_get_binning_layout(capacity, C.byref(bl))
        if len(_layout_cache) > 256:
            _layout_cache.clear()
        v = _layout_cache[capacity] = int(bl.total)
    return v


def _align(n):
    return (n + 511) & ~511


try:                                       # raw cudaStream_t of the current stream without building a torch.cuda.Stream object
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                     # older / newer torch: the public, slower route
    _raw_stream = None


def _current_stream(dev):
    if _raw_stream is not None and dev.index is not None:
        return _raw_stream(dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


def _select_device(L, dev):
    # unconditionally: torch.cuda.set_device / another engine may have changed the thread's device since the last call,
    # and cudaSetDevice on the current device is free
    if dev.index is not None:
        L.lvdgs_set_device(dev.index)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta,
                        rho, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, theta, rho, raster_settings)


def _params(rs: GaussianRasterizationSettings, P: int, M: int) -> _native.RasterParams:
    return _native.RasterParams(P, int(rs.sh_degree), M, int(rs.image_width), int(rs.image_height), float(rs.tanfovx),
                                float(rs.tanfovy), float(rs.scale_modifier), int(bool(rs.prefiltered)),
                                int(bool(rs.debug)), FLAGS)


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
        bg = _prep(_on(rs.bg, dev)); view = _prep(_on(rs.viewmatrix, dev)); proj = _prep(_on(rs.projmatrix, dev))
        praw = _prep(_on(rs.projmatrix_raw, dev)); campos = _prep(_on(rs.campos, dev))
        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        opac_img = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        n_touched = torch.empty((P,), dtype=torch.int32, device=dev)
        bufs = {}
        _tls.store, _tls.dev = bufs, dev
        prm = _params(rs, P, M)
        R = C.c_int64(0)
        cap = C.c_int64(0)
        hint = _capacity_hint.get(dev.index, 0) if SPECULATIVE else 0
        _select_device(L, dev)
        stream = _current_stream(dev)
        p = _ptr
        # one allocation for the geometry / image / binning state (binning only when a capacity hint sizes it)
        global _static_cb
        if _static_cb is None:
            _static_cb = C.cast(L.lvdgs_static_resize, _native.RESIZE_FN)
        g_bytes, i_bytes = _layout_sizes(L, P, W, H) if P > 0 else (0, _layout_sizes(L, 0, W, H)[1])
        b_bytes = _binning_size(L, hint) if (hint > 0 and P > 0) else 0
        o_img, o_bin = _align(g_bytes), _align(g_bytes) + _align(i_bytes)
        arena = torch.empty(o_bin + _align(b_bytes), dtype=torch.uint8, device=dev)
        a0 = arena.data_ptr()
        sb = _native.StaticBuffers()
        sb.base[0], sb.base[1], sb.base[2] = a0 if g_bytes else None, (a0 + o_bin) if b_bytes else None, a0 + o_img
        sb.capacity[0], sb.capacity[1], sb.capacity[2] = g_bytes, b_bytes, i_bytes
        sb.fallback, sb.fallback_user = _RESIZE_CB, None
        try:
            rc = L.lvdgs_rasterize_forward(C.byref(prm), p(bg), p(m3), p(cp), p(op), p(sc), p(rot), p(cov), p(view),
                                           p(proj), p(praw), p(shs), p(campos), _static_cb, C.addressof(sb), hint, p(color),
                                           p(radii), p(depth), p(opac_img), p(n_touched), C.byref(R), C.byref(cap), stream)
        finally:
            _tls.store = None
        _native.check(rc, "lvdgs_rasterize_forward")
        ctx.rs = rs
        ctx.flags = prm.flags          # the backward must see the forward's sort layout
        ctx.num_rendered = R.value
        ctx.capacity = cap.value
        if SPECULATIVE:   # grow at once; shrink (slowly) only when the hint is far above what the views need, so that the
            want = int(R.value * 1.25) + 65536          # binning size -- and with it the cached layout -- is stable frame to frame
            _capacity_hint[dev.index] = want if want > hint else (int(hint * 0.98) if want * 2 < hint else hint)
        # live pointers of the three buffers: the arena's parts, or what the fallback callback allocated for an oversized request
        ctx.buf_ptrs = tuple(bufs[w].data_ptr() if w in bufs else sb.base[w] for w in (0, 1, 2))
        ctx.arena = (arena, (0, g_bytes), (o_bin, b_bytes), (o_img, i_bytes), bufs)      # keeps the memory alive until the backward has run
        ctx.prm = prm
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
        # tracking differentiates w.r.t. the camera only (the frontend's Gaussians are detached copies,
        # utils/multiprocessing_utils.py:29-30): then no per-Gaussian parameter gradient is computed or stored
        nig = ctx.needs_input_grad
        pose_only = not (nig[0] or nig[2] or nig[3] or nig[4] or nig[5] or nig[6] or nig[7])
        # one allocation for every gradient tensor (+ the backward scratch), handed out as views
        widths = [("means2D", 3)]
        if not pose_only:
            widths += [("opac", 1), ("means3D", 3)]
            if cp is not None: widths.append(("colors", 3))
            if cov is not None: widths.append(("cov", 6))
            else: widths += [("sc", 3), ("rot", 4)]
            if shs is not None: widths.append(("sh", 3 * M))
        nscratch = (L.lvdgs_backward_scratch_bytes(P, ctx.num_rendered) + 3) // 4
        pad4 = lambda n: (n + 3) & ~3                  # every view starts 16-byte aligned (float4 stores in the kernels)
        ngrad = 8 + sum(pad4(w * P) for _, w in widths)
        flat = torch.empty((ngrad + nscratch,), dtype=torch.float32, device=dev)
        if ZEROED_OUTPUTS:
            flat[:ngrad].zero_()                      # LVDGS_FLAG_ZEROED_OUTPUTS: the backward then visits visible Gaussians only
        g, off = {}, 8                                # first 8 floats: dL_dtau_sum
        for name, w in widths:
            g[name] = flat[off:off + w * P]
            off += pad4(w * P)
        flat_s = flat[off:off + nscratch]
        g_tau = flat[0:6]
        prm = ctx.prm                                          # the forward's parameter block (same shapes, same camera)
        prm.flags = ctx.flags | (32 if ZEROED_OUTPUTS else 0)   # LVDGS_FLAG_ZEROED_OUTPUTS
        if pose_only:
            prm.flags |= 8          # LVDGS_FLAG_POSE_ONLY
        _select_device(L, dev)
        stream = _current_stream(dev)
        p = _ptr
        gc = _prep(grad_out_color)
        gd = _prep(grad_out_depth) if grad_out_depth is not None else None
        go = _prep(grad_out_opacity) if (grad_out_opacity is not None and (FLAGS & 2)) else None
        b0, b1, b2 = ctx.buf_ptrs
        rc = L.lvdgs_rasterize_backward(C.byref(prm), p(bg), p(m3), p(radii), p(cp), p(op), p(sc), p(rot), p(cov), p(view),
                                        p(proj), p(praw), p(gc), p(gd), p(go), p(shs), p(campos), b0,
                                        ctx.num_rendered, ctx.capacity, b1, b2, p(flat_s),
                                        flat_s.numel() * 4, p(g["means2D"]), p(g.get("colors")), p(g.get("opac")),
                                        p(g.get("means3D")), p(g.get("cov")), p(g.get("sh")), p(g.get("sc")), p(g.get("rot")),
                                        None, p(g_tau), stream)
        _native.check(rc, "lvdgs_rasterize_backward")
        th_shape, rho_shape = ctx.in_shapes
        g_rho = g_tau[:3].reshape(rho_shape) if rho_shape is not None else None
        g_theta = g_tau[3:].reshape(th_shape) if th_shape is not None else None
        v = lambda name, *shape: g[name].view(*shape) if name in g else None
        return (v("means3D", P, 3), v("means2D", P, 3), v("sh", P, M, 3), v("colors", P, 3), v("opac", P, 1), v("sc", P, 3),
                v("rot", P, 4), v("cov", P, 6), g_theta, g_rho, None)


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
