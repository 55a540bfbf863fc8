"""Seeded synthetic Gaussian clouds and pinhole cameras of the shapes BASELINE.json names.

Restates the two helpers of the missing `gaussian_splatting/utils/graphics_utils.py` that the reference
uses to build the rasterizer's matrices (call sites: /root/reference/utils/slam_frontend.py:1743-1748,
utils/camera_utils.py:90-92,106-120): `getProjectionMatrix2` and `getWorld2View2`.  numpy only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# name -> (W, H, fx, fy, cx, cy); values from /root/reference/configs/mono/KITTI/00.yaml:8-18 and 08.yaml:8-18
CAMERAS = {
    "kitti": (1241, 376, 718.856, 718.856, 607.1928, 185.2157),
    "nuscenes": (1600, 900, 707.0912, 707.0912, 601.8873, 183.1104),
    "hd": (1920, 1080, 0.6 * 1920, 0.6 * 1920, 1920 / 2 - 7.3, 1080 / 2 + 3.1),
    "vga": (640, 480, 0.6 * 640, 0.6 * 640, 640 / 2 - 7.3, 480 / 2 + 3.1),
    "mast3r_kitti": (512, 144, 718.856 * 512 / 1241, 718.856 * 144 / 376, 607.1928 * 512 / 1241, 185.2157 * 144 / 376),
}


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def getWorld2View2(R, t):
    Rt = np.eye(4, dtype=np.float64)
    Rt[:3, :3] = R
    Rt[:3, 3] = t
    return Rt


def getProjectionMatrix2(znear, zfar, cx, cy, fx, fy, W, H):
    left = ((2 * cx - W) / W - 1.0) * W / 2.0
    right = ((2 * cx - W) / W + 1.0) * W / 2.0
    top = ((2 * cy - H) / H + 1.0) * H / 2.0
    bottom = ((2 * cy - H) / H - 1.0) * H / 2.0
    left, right = znear / fx * left, znear / fx * right
    top, bottom = znear / fy * top, znear / fy * bottom
    P = np.zeros((4, 4), dtype=np.float64)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class Cam:
    """The fields of utils/camera_utils.py:Camera that the renderer reads, as float32 numpy arrays."""
    image_width: int
    image_height: int
    fx: float
    fy: float
    cx: float
    cy: float
    R: np.ndarray          # world->camera rotation (3,3)
    T: np.ndarray          # world->camera translation (3,)

    @property
    def FoVx(self):
        return focal2fov(self.fx, self.image_width)

    @property
    def FoVy(self):
        return focal2fov(self.fy, self.image_height)

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)

    @property
    def projection_matrix(self):   # = getProjectionMatrix2(...).transpose(0,1)
        return getProjectionMatrix2(0.01, 100.0, self.cx, self.cy, self.fx, self.fy, self.image_width,
                                    self.image_height).T.astype(np.float32)

    @property
    def world_view_transform(self):
        return getWorld2View2(self.R, self.T).T.astype(np.float32)

    @property
    def full_proj_transform(self):
        return (self.world_view_transform.astype(np.float64) @ self.projection_matrix.astype(np.float64)).astype(np.float32)

    @property
    def camera_center(self):
        return np.linalg.inv(self.world_view_transform.astype(np.float64))[3, :3].astype(np.float32)


def make_camera(name="kitti", k=None, centered=False):
    """Camera `k` of a keyframe window: translate 0.8*k m along +z, yaw (k-3.5) degrees. k=None -> identity."""
    W, H, fx, fy, cx, cy = CAMERAS[name]
    if centered:
        cx, cy = W / 2.0, H / 2.0
    R = np.eye(3)
    T = np.zeros(3)
    if k is not None:
        a = math.radians(k - 3.5)
        R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
        c2w_t = np.array([0.0, 0.0, 0.8 * k])
        T = -R @ c2w_t
    return Cam(W, H, fx, fy, cx, cy, R, T)


def make_scene(N, cam: Cam, seed=0, sh_degree=0, behind_frac=0.02):
    """Street-scene-like frustum fill. Returns dict of float32 arrays in the rasterizer's input layout."""
    rng = np.random.default_rng(seed)
    z = np.exp(rng.uniform(math.log(1.5), math.log(80.0), N))
    nb = int(N * behind_frac)
    if nb:
        z[:nb] = rng.uniform(-5.0, 0.2, nb)
    zz = np.abs(z) + 0.3
    x = zz * cam.tanfovx * rng.uniform(-1.25, 1.25, N)
    y = zz * cam.tanfovy * rng.uniform(-1.25, 1.25, N)
    means = np.stack([x, y, z], 1)
    perm = rng.permutation(N)
    means = means[perm]
    zz = zz[perm]
    base = (zz / cam.fx) * np.exp(rng.uniform(math.log(0.7), math.log(6.0), N))
    scales = base[:, None] * np.exp(rng.normal(0.0, 0.35, (N, 3)))
    rots = rng.normal(0, 1, (N, 4))
    rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    opac = rng.uniform(0.02, 0.98, (N, 1))
    M = (sh_degree + 1) ** 2
    shs = rng.uniform(-2.2, 2.2, (N, M, 3))   # ~10% of SH0 channels hit the max(0,.) clamp
    if M > 1:
        shs[:, 1:] *= 0.3
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(means3D=f32(means), scales=f32(scales), rotations=f32(rots), opacities=f32(opac), shs=f32(shs),
                sh_degree=sh_degree)


def make_upstream_grads(cam: Cam, seed=1):
    """grad_color = N(0,1)[3,H,W]/(H*W), grad_depth = N(0,1)[1,H,W]/(H*W)."""
    rng = np.random.default_rng(seed)
    H, W = cam.image_height, cam.image_width
    gc = (rng.normal(0, 1, (3, H, W)) / (H * W)).astype(np.float32)
    gd = (rng.normal(0, 1, (1, H, W)) / (H * W)).astype(np.float32)
    return gc, gd
