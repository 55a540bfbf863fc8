"""Host side of the callers either side of the rasterizer, over the C ABI.

* `get_loss_tracking`, `get_loss_mapping` (+ `_rgb` / `_rgbd` variants): same names, arguments and values as
  /root/reference/utils/slam_utils.py:42-121, computed -- loss and gradient -- by ONE kernel (`lvdgs_fused_loss`)
  instead of ~10 torch elementwise kernels forward and as many backward.
* `covisibility(a, b)`, `accumulate_n_obs(masks)`: the keyframe-management counts of utils/slam_frontend.py:1598-1643
  and utils/slam_backend.py:322-325 without temporaries or host copies.
* `compact_rows(keep, tensors)`: the boolean-mask indexing of every parameter / Adam-moment tensor that
  GaussianModel.prune_points performs (utils/slam_backend.py:128-145,322-339), for all tensors in one launch.
* `gather_rows(index, tensors, out)`: the row copies of densify_and_clone / densify_and_split (utils/slam_backend.py:359-376),
  for all tensors in one launch, written where the caller wants them (e.g. the tail of a larger buffer).

CUDA only: there is no CPU path (a CPU tensor raises).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native

LOSS_OPACITY_WEIGHT, LOSS_DEPTH_NEEDS_OPAQUE = 1, 2
_ws = {}


def _workspace(dev, nbytes):
    # one workspace per (device, stream): the loss kernel keeps per-block partials and a ticket in it, so two launches
    # that may run concurrently must not share one
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(int(nbytes), dtype=torch.uint8, device=dev)      # zeroed: holds the loss kernel's ticket
        _ws[key] = buf
    return buf


def _f32(t):
    return t if (t.dtype is torch.float32 and t.is_contiguous()) else t.contiguous().float()


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"lvdgs.slam_ops.{what}: tensors must be on a CUDA device; there is no CPU path")


class _FusedLoss(torch.autograd.Function):
    """loss, with dL/d(image, depth, opacity, exposure_a, exposure_b) produced by the same kernel."""

    @staticmethod
    def forward(ctx, image, depth, opacity, exposure_a, exposure_b, gt_image, gt_depth, grad_mask, thr, w_rgb, w_depth, flags):
        _need_cuda(image, "loss")
        L = _native.lib()
        dev = image.device
        _, H, W = image.shape
        img = _f32(image)
        dep = None if depth is None else _f32(depth)
        opa = None if opacity is None else _f32(opacity)
        gt = _f32(gt_image)
        gtd = None if gt_depth is None else _f32(gt_depth)
        gm = None if grad_mask is None else _f32(grad_mask)
        expo = None
        if exposure_a is not None:
            expo = torch.cat([exposure_a.detach().reshape(1).float(), exposure_b.detach().reshape(1).float()])
        g_img = torch.empty_like(img)
        g_dep = torch.empty_like(dep) if dep is not None else None
        g_opa = torch.empty_like(opa) if (opa is not None and ctx.needs_input_grad[2]) else None
        out = torch.empty(4, dtype=torch.float32, device=dev)
        ws = _workspace(dev, L.lvdgs_fused_loss_workspace_bytes())
        p = _native.ptr
        rc = L.lvdgs_fused_loss(W, H, p(img), p(dep), p(opa), p(gt), p(gtd), p(gm), p(expo), float(thr), float(w_rgb),
                                float(w_depth), int(flags), p(g_img), p(g_dep), p(g_opa), p(out), p(ws), ws.numel(),
                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _native.check(rc, "lvdgs_fused_loss")
        ctx.save_for_backward(g_img, g_dep, g_opa, out)
        ctx.shapes = (image.shape, None if depth is None else depth.shape, None if opacity is None else opacity.shape,
                      None if exposure_a is None else exposure_a.shape, None if exposure_b is None else exposure_b.shape)
        return out[0]

    @staticmethod
    def backward(ctx, go):
        g_img, g_dep, g_opa, out = ctx.saved_tensors
        s_img, s_dep, s_opa, s_ea, s_eb = ctx.shapes
        nig = ctx.needs_input_grad
        gi = (g_img * go).view(s_img) if nig[0] else None
        gd = (g_dep * go).view(s_dep) if (nig[1] and g_dep is not None) else None
        gop = (g_opa * go).view(s_opa) if (nig[2] and g_opa is not None) else None
        gea = (out[1] * go).reshape(s_ea) if (nig[3] and s_ea is not None) else None
        geb = (out[2] * go).reshape(s_eb) if (nig[4] and s_eb is not None) else None
        return gi, gd, gop, gea, geb, None, None, None, None, None, None, None


def fused_loss_into(image, depth, gt_image, gt_depth, g_image, g_depth, out, *, opacity=None, grad_mask=None, exposure=None,
                    rgb_boundary_threshold=0.01, w_rgb=1.0, w_depth=0.0, flags=0):
    """The loss kernel without autograd, for allocation-free loops over the C ABI (lvdgs.engine.RasterEngine.run_views'
    `upstream` hook): writes dL/dimage into g_image, dL/ddepth into g_depth and {loss, dL/da, dL/db, 0} into out[4] --
    all caller-owned float32 CUDA tensors."""
    L = _native.lib()
    dev = image.device
    _, H, W = image.shape
    ws = _workspace(dev, L.lvdgs_fused_loss_workspace_bytes())
    p = _native.ptr
    rc = L.lvdgs_fused_loss(W, H, p(image), p(depth), p(opacity), p(gt_image), p(gt_depth), p(grad_mask), p(exposure),
                            float(rgb_boundary_threshold), float(w_rgb), float(w_depth), int(flags), p(g_image), p(g_depth), None,
                            p(out), p(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _native.check(rc, "lvdgs_fused_loss")
    return out


def fused_loss(image, depth=None, opacity=None, exposure_a=None, exposure_b=None, *, gt_image, gt_depth=None,
               grad_mask=None, rgb_boundary_threshold=0.01, w_rgb=1.0, w_depth=0.0, flags=0):
    return _FusedLoss.apply(image, depth, opacity, exposure_a, exposure_b, gt_image, gt_depth, grad_mask,
                            rgb_boundary_threshold, w_rgb, w_depth, flags)


def _gt_depth(viewpoint, device):
    md = viewpoint.mono_depth
    if not torch.is_tensor(md):
        md = torch.from_numpy(md)
    return md.to(dtype=torch.float32, device=device)[None]


# ---- the reference's loss entry points (utils/slam_utils.py:42-121), same signatures ----
def get_loss_tracking(config, image, depth, opacity, viewpoint, initialization=False):
    if config["Training"]["monocular"]:
        return get_loss_tracking_rgb(config, image, depth, opacity, viewpoint, _exposure=True)
    return get_loss_tracking_rgbd(config, image, depth, opacity, viewpoint, _exposure=True)


def get_loss_tracking_rgb(config, image, depth, opacity, viewpoint, _exposure=False):
    ea, eb = (viewpoint.exposure_a, viewpoint.exposure_b) if _exposure else (None, None)
    return fused_loss(image, None, opacity, ea, eb, gt_image=viewpoint.original_image.to(image.device),
                      grad_mask=viewpoint.grad_mask, rgb_boundary_threshold=config["Training"]["rgb_boundary_threshold"],
                      w_rgb=1.0, w_depth=0.0, flags=LOSS_OPACITY_WEIGHT)


def get_loss_tracking_rgbd(config, image, depth, opacity, viewpoint, initialization=False, _exposure=False):
    alpha = config["Training"]["alpha"] if "alpha" in config["Training"] else 0.95
    ea, eb = (viewpoint.exposure_a, viewpoint.exposure_b) if _exposure else (None, None)
    return fused_loss(image, depth, opacity, ea, eb, gt_image=viewpoint.original_image.to(image.device),
                      gt_depth=_gt_depth(viewpoint, image.device), grad_mask=viewpoint.grad_mask,
                      rgb_boundary_threshold=config["Training"]["rgb_boundary_threshold"], w_rgb=alpha, w_depth=1 - alpha,
                      flags=LOSS_OPACITY_WEIGHT | LOSS_DEPTH_NEEDS_OPAQUE)


def get_loss_mapping(config, image, viewpoint, depth=None, initialization=False, monodepth=True):
    ea, eb = (None, None) if initialization else (viewpoint.exposure_a, viewpoint.exposure_b)
    if config["Training"]["monocular"] and not monodepth:
        return get_loss_mapping_rgb(config, image, viewpoint, _exposure=(ea, eb))
    return get_loss_mapping_rgbd(config, image, depth, viewpoint, _exposure=(ea, eb))


def get_loss_mapping_rgb(config, image, viewpoint, _exposure=(None, None)):
    return fused_loss(image, None, None, *_exposure, gt_image=viewpoint.original_image.to(image.device),
                      rgb_boundary_threshold=config["Training"]["rgb_boundary_threshold"], w_rgb=1.0, w_depth=0.0)


def get_loss_mapping_rgbd(config, image, depth, viewpoint, initialization=False, _exposure=(None, None)):
    alpha = config["Training"]["alpha"] if "alpha" in config["Training"] else 0.95
    return fused_loss(image, depth, None, *_exposure, gt_image=viewpoint.original_image.to(image.device),
                      gt_depth=_gt_depth(viewpoint, image.device),
                      rgb_boundary_threshold=config["Training"]["rgb_boundary_threshold"], w_rgb=alpha, w_depth=1 - alpha)


# ---- the masked L1 + SSIM + depth mapping loss (utils/slam_backend.py:199-261) ----
class _MaskedMappingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, gt_image, static_mask, background, mono_depth, lambda_dssim, depth_lambda):
        _need_cuda(image, "masked_mapping_loss")
        L = _native.lib()
        dev = image.device
        _, H, W = image.shape
        img, gt = _f32(image), _f32(gt_image)
        bg = _f32(background.to(dev))
        mask = None if static_mask is None else (static_mask.to(dev) != 0).reshape(H, W).contiguous().view(torch.uint8)
        dep = None if depth is None else _f32(depth).reshape(H, W)
        mono = None if (mono_depth is None or depth is None) else _f32(mono_depth.to(dev)).reshape(H, W)
        g_img = torch.empty_like(img)
        g_dep = torch.empty_like(dep) if dep is not None else None
        out = torch.empty(8, dtype=torch.float32, device=dev)
        nbytes = L.lvdgs_masked_ssim_loss_workspace_bytes(W, H)
        key = ("ssim", dev.index, torch.cuda.current_stream(dev).cuda_stream, W, H)
        ws = _ws.get(key)
        if ws is None:
            ws = _ws[key] = torch.zeros(int(nbytes), dtype=torch.uint8, device=dev)
        p = _native.ptr
        rc = L.lvdgs_masked_ssim_loss(W, H, p(img), p(gt), p(mask), p(bg), p(dep), p(mono), float(lambda_dssim), float(depth_lambda),
                                      p(g_img), p(g_dep), p(out), p(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _native.check(rc, "lvdgs_masked_ssim_loss")
        ctx.save_for_backward(g_img, g_dep)
        ctx.shapes = (image.shape, None if depth is None else depth.shape)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, go, _gout):
        g_img, g_dep = ctx.saved_tensors
        s_img, s_dep = ctx.shapes
        gi = (g_img * go).view(s_img) if ctx.needs_input_grad[0] else None
        gd = (g_dep * go).view(s_dep) if (ctx.needs_input_grad[1] and g_dep is not None) else None
        return gi, gd, None, None, None, None, None, None


def masked_mapping_loss(image, gt_image, static_mask, background, depth=None, mono_depth=None, lambda_dssim=0.2,
                        depth_lambda=0.1, return_terms=False):
    """The per-keyframe loss of utils/slam_backend.py:199-261 (dynamic pixels painted with the background in render and
    ground truth; (1 - lambda) L1 + lambda (1 - SSIM); + depth_lambda * masked mean |depth - mono_depth|) -- value and
    gradients from two tiled kernels (lvdgs_masked_ssim_loss).  static_mask: bool [H,W], True = static (kept) pixel."""
    loss, terms = _MaskedMappingLoss.apply(image, depth, gt_image, static_mask, background, mono_depth, lambda_dssim, depth_lambda)
    return (loss, terms) if return_terms else loss


# ---- covisibility ----
def _mask_arg(t):
    _need_cuda(t, "covisibility")
    t = t.contiguous()
    if t.dtype is torch.bool:
        t = t.view(torch.uint8)
    if t.element_size() not in (1, 4, 8) or t.dtype.is_floating_point:
        t = (t != 0).view(torch.uint8)
    return t


def covisibility(a, b):
    """-> int64 tensor [4] on the device: |a|, |b|, |a and b|, |a or b| (elements count when non-zero)."""
    a, b = _mask_arg(a), _mask_arg(b)
    if a.shape != b.shape or a.element_size() != b.element_size():
        raise ValueError("covisibility: masks must have the same shape and element size")
    L = _native.lib()
    out = torch.empty(4, dtype=torch.int64, device=a.device)
    rc = L.lvdgs_covis_counts(a.numel(), _native.ptr(a), _native.ptr(b), a.element_size(), _native.ptr(out),
                              C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    _native.check(rc, "lvdgs_covis_counts")
    return out


def accumulate_n_obs(masks):
    """n_obs[i] = number of visibility arrays that see Gaussian i (int32, on the device)."""
    ms = [_mask_arg(m) for m in masks]
    if not ms:
        raise ValueError("accumulate_n_obs: no masks")
    n, es, dev = ms[0].numel(), ms[0].element_size(), ms[0].device
    if any(m.numel() != n or m.element_size() != es for m in ms):
        raise ValueError("accumulate_n_obs: masks must have the same length and element size")
    L = _native.lib()
    ptrs = torch.tensor([m.data_ptr() for m in ms], dtype=torch.int64).to(dev)
    out = torch.empty(n, dtype=torch.int32, device=dev)
    rc = L.lvdgs_n_obs(n, len(ms), _native.ptr(ptrs), es, _native.ptr(out),
                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _native.check(rc, "lvdgs_n_obs")
    return out


# ---- prune ----
def compact_rows(keep, tensors, out=None):
    """Rows of every tensor in `tensors` (each [n, ...], float32) where `keep` is non-zero, in order -- what
    `t[keep]` gives for each, from one scan of the mask and one launch per 16 tensors.  Returns the list of new tensors
    (`out`: optional preallocated destinations with at least as many rows; views of the first `count` rows are returned)."""
    _need_cuda(keep, "compact_rows")
    L = _native.lib()
    dev = keep.device
    k8 = keep.contiguous()
    k8 = k8.view(torch.uint8) if k8.dtype is torch.bool else (k8 != 0).view(torch.uint8)
    n = k8.numel()
    srcs = [_f32(t) for t in tensors]
    for t in srcs:
        if t.shape[0] != n:
            raise ValueError("compact_rows: every tensor needs one row per mask element")
    ws = torch.empty(L.lvdgs_compact_workspace_bytes(n), dtype=torch.uint8, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    cnt_ptr = C.c_void_p()
    _native.check(L.lvdgs_compact_count(n, _native.ptr(k8), _native.ptr(ws), ws.numel(), C.byref(cnt_ptr), stream),
                  "lvdgs_compact_count")
    off = cnt_ptr.value - ws.data_ptr()
    count = int(ws[off:off + 4].view(torch.int32).item())            # the one host read-back: the new map size
    dsts = []
    for i, t in enumerate(srcs):
        if out is not None:
            d = out[i]
            if d.shape[0] < count or d.data_ptr() == t.data_ptr():
                raise ValueError("compact_rows: destination too small or aliasing its source")
            dsts.append(d[:count])
        else:
            dsts.append(torch.empty((count,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev))
    if count == 0:
        return dsts
    for lo in range(0, len(srcs), 16):
        chunk = list(range(lo, min(lo + 16, len(srcs))))
        m = len(chunk)
        sp = (C.c_void_p * m)(*[srcs[i].data_ptr() for i in chunk])
        dp = (C.c_void_p * m)(*[dsts[i].data_ptr() for i in chunk])
        wd = (C.c_int32 * m)(*[max(1, srcs[i][0].numel()) if n else 1 for i in chunk])
        _native.check(L.lvdgs_compact_move(n, _native.ptr(k8), _native.ptr(ws), m, sp, dp, wd, stream), "lvdgs_compact_move")
    return dsts


# ---- densify ----
def gather_rows(index, tensors, out=None):
    """out[k][j] = tensors[k][index[j]] for every tensor (each [n, ...], float32), one launch per 16 tensors -- what
    `t[index]` gives for each.  `out`: optional destinations with at least len(index) rows (e.g. views of the tail of a
    preallocated parameter buffer); returns the list of filled [len(index), ...] tensors."""
    _need_cuda(index, "gather_rows")
    L = _native.lib()
    dev = index.device
    idx = index.contiguous().to(torch.int64)
    m = idx.numel()
    srcs = [_f32(t) for t in tensors]
    n = srcs[0].shape[0] if srcs else 0
    if any(t.shape[0] != n for t in srcs):
        raise ValueError("gather_rows: every tensor needs the same number of rows")
    dsts = []
    for i, t in enumerate(srcs):
        d = out[i][:m] if out is not None else torch.empty((m,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
        if d.shape[0] < m or not d.is_contiguous():
            raise ValueError("gather_rows: destination too small or not contiguous")
        dsts.append(d)
    if m == 0 or n == 0:
        return dsts
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for lo in range(0, len(srcs), 16):
        chunk = list(range(lo, min(lo + 16, len(srcs))))
        k = len(chunk)
        sp = (C.c_void_p * k)(*[srcs[i].data_ptr() for i in chunk])
        dp = (C.c_void_p * k)(*[dsts[i].data_ptr() for i in chunk])
        wd = (C.c_int32 * k)(*[max(1, srcs[i][0].numel()) for i in chunk])
        _native.check(L.lvdgs_gather_rows(m, _native.ptr(idx), n, k, sp, dp, wd, stream), "lvdgs_gather_rows")
    return dsts
