"""`gaussian_splatting.gaussian_renderer.render` / `render_with_custom_resolution` -- the renderer glue the
reference imports but does not ship.  Signatures and the returned dict keys
are pinned by the call sites: utils/slam_backend.py:98-117,184-194,277-296,407-414; utils/slam_frontend.py:1493-1500;
utils/eval_utils_0806.py:215-219; utils/init_pose.py:145-146.
"""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def _render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, mask, H, W):
    if pc.get_xyz.shape[0] == 0:
        return None
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device=pc.get_xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(H), image_width=int(W), tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    means3D = pc.get_xyz
    means2D = screenspace_points
    opacity = pc.get_opacity
    scales = rotations = cov3D_precomp = None
    if getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales = pc.get_scaling
        rotations = pc.get_rotation
    shs = colors_precomp = None
    if override_color is not None:
        colors_precomp = override_color
    else:
        shs = pc.get_features
    theta = getattr(viewpoint_camera, "cam_rot_delta", None)
    rho = getattr(viewpoint_camera, "cam_trans_delta", None)
    if mask is not None:
        sel = lambda t: None if t is None else t[mask]
        means3D, means2D, opacity = means3D[mask], means2D[mask], opacity[mask]
        shs, colors_precomp, scales, rotations, cov3D_precomp = sel(shs), sel(colors_precomp), sel(scales), sel(rotations), sel(cov3D_precomp)
    rendered_image, radii, depth, opacity_img, n_touched = rasterizer(
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity, scales=scales,
        rotations=rotations, cov3D_precomp=cov3D_precomp, theta=theta, rho=rho)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii, "depth": depth, "opacity": opacity_img, "n_touched": n_touched}


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, mask=None):
    if pc.get_xyz.shape[0] == 0:
        return None
    return _render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, mask,
                   viewpoint_camera.image_height, viewpoint_camera.image_width)


def render_with_custom_resolution(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None,
                                  mask=None, target_width=None, target_height=None):
    if pc.get_xyz.shape[0] == 0:
        return None
    W = target_width if target_width is not None else viewpoint_camera.image_width
    H = target_height if target_height is not None else viewpoint_camera.image_height
    return _render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, mask, H, W)
