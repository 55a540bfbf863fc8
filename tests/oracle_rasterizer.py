"""A `diff_gaussian_rasterization`-shaped module backed by the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Lets the reference's own Python (utils.camera_utils.Camera -> gaussian_renderer.render -> utils.slam_utils losses ->
loss.backward() -> utils.pose_utils.update_pose) run end to end on a box without a GPU: `install()` registers this
module as `diff_gaussian_rasterization`, so the shim's `render` builds its settings and calls the rasterizer exactly as
it does on the product path, and the oracle (oracle/raster_oracle.c) does the arithmetic.  Used by
tests/golden/make_reference_golden.py to record golden tracking trajectories and by tests/test_reference_pin.py.
The product package never imports this file.
"""
from __future__ import annotations

import sys
from typing import NamedTuple

import numpy as np
import torch

import oracle

FLAGS = 0       # oracle.FLAG_EXACT_PP | oracle.FLAG_OPACITY_GRAD for the true derivative; 0 = upstream behaviour


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


def _np(t):
    return None if t is None or t.numel() == 0 else t.detach().cpu().numpy().astype(np.float32)


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho, rs):
        fwd = oracle.rasterize_forward(_np(means3D), _np(opacities), _np(scales), _np(rotations), _np(sh), _np(colors_precomp),
                                       _np(cov3Ds_precomp), viewmatrix=_np(rs.viewmatrix), projmatrix=_np(rs.projmatrix),
                                       campos=_np(rs.campos), bg=_np(rs.bg), W=int(rs.image_width), H=int(rs.image_height),
                                       tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), sh_degree=int(rs.sh_degree),
                                       scale_modifier=float(rs.scale_modifier), want_margin=False)
        ctx.fwd, ctx.rs = fwd, rs
        ctx.shapes = (means3D.shape, None if sh.numel() == 0 else sh.shape, theta.shape if theta.numel() else None,
                      rho.shape if rho.numel() else None)
        t = torch.from_numpy
        radii, n_touched = t(fwd["radii"].copy()), t(fwd["n_touched"].copy())
        ctx.mark_non_differentiable(radii, n_touched)
        return t(fwd["color"].copy()), radii, t(fwd["depth"].copy()), t(fwd["opacity"].copy()), n_touched

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_opacity, g_nt):
        fwd, rs = ctx.fwd, ctx.rs
        go = _np(g_opacity) if (FLAGS & oracle.FLAG_OPACITY_GRAD) and g_opacity is not None else None
        g = oracle.rasterize_backward(fwd, _np(g_color), None if g_depth is None else _np(g_depth), go,
                                      projmatrix_raw=_np(rs.projmatrix_raw), flags=FLAGS)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        P = fwd["P"]
        m2d = np.zeros((P, 3), np.float32)
        m2d[:, :2] = g["dL_dmean2D"]
        mshape, shshape, th_shape, rho_shape = ctx.shapes
        g_theta = t(g["grad_theta"]).reshape(th_shape) if th_shape is not None else None
        g_rho = t(g["grad_rho"]).reshape(rho_shape) if rho_shape is not None else None
        return (t(g["dL_dmeans3D"]), t(m2d), t(g["dL_dsh"]), t(g["dL_dcolors_precomp"]), t(g["dL_dopacity"]).reshape(P, 1),
                t(g["dL_dscales"]), t(g["dL_drots"]), t(g["dL_dcov3D"]) if fwd["_in"]["cov3D_precomp"] is not None else None,
                g_theta, g_rho, None)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        e = torch.Tensor([])
        f = lambda x: e if x is None else x
        return _Rasterize.apply(means3D, means2D, f(shs), f(colors_precomp), opacities, f(scales), f(rotations),
                                f(cov3D_precomp), f(theta), f(rho), self.raster_settings)


def install():
    """Register this module as `diff_gaussian_rasterization` (before the shim's gaussian_renderer is imported)."""
    sys.modules["diff_gaussian_rasterization"] = sys.modules[__name__]
    sys.modules.pop("gaussian_splatting.gaussian_renderer", None)
