"""The reference-facing call path end to end: gaussian_splatting.gaussian_renderer.render with a Camera-like object
and a GaussianModel-like object, exactly as utils/slam_frontend.py:1493-1521 uses it -- including a short tracking loop
(pose-only optimisation through theta/rho gradients + the reference's update_pose convention)."""
import math

import numpy as np
import pytest
import torch

from lvdgs import synth
from gaussian_splatting.gaussian_renderer import render, render_with_custom_resolution

pytestmark = pytest.mark.gpu


# ---- restated from /root/reference/utils/pose_utils.py:11-87 (the consumer of the pose gradients) ----
def skew(x):
    s = torch.zeros(3, 3, device=x.device, dtype=x.dtype)
    s[0, 1], s[0, 2], s[1, 0], s[1, 2], s[2, 0], s[2, 1] = -x[2], x[1], x[2], -x[0], -x[1], x[0]
    return s


def SE3_exp(tau):
    rho, theta = tau[:3], tau[3:]
    W = skew(theta); W2 = W @ W
    ang = torch.norm(theta)
    I = torch.eye(3, device=tau.device, dtype=tau.dtype)
    if ang < 1e-5:
        R = I + W + 0.5 * W2; V = I + 0.5 * W + W2 / 6.0
    else:
        R = I + (torch.sin(ang) / ang) * W + ((1 - torch.cos(ang)) / ang ** 2) * W2
        V = I + W * ((1 - torch.cos(ang)) / ang ** 2) + W2 * ((ang - torch.sin(ang)) / ang ** 3)
    T = torch.eye(4, device=tau.device, dtype=tau.dtype)
    T[:3, :3] = R; T[:3, 3] = V @ rho
    return T


class Cam(torch.nn.Module):
    """Fields of utils/camera_utils.py:Camera that render() reads (:42-56,106-120)."""

    def __init__(self, c: synth.Cam, dev):
        super().__init__()
        self.R = torch.tensor(c.R, dtype=torch.float32, device=dev)
        self.T = torch.tensor(c.T, dtype=torch.float32, device=dev)
        self.FoVx, self.FoVy = c.FoVx, c.FoVy
        self.image_height, self.image_width = c.image_height, c.image_width
        self.projection_matrix = torch.tensor(c.projection_matrix, device=dev)
        self.cam_rot_delta = torch.nn.Parameter(torch.zeros(3, device=dev))
        self.cam_trans_delta = torch.nn.Parameter(torch.zeros(3, device=dev))

    @property
    def world_view_transform(self):
        Rt = torch.eye(4, device=self.R.device)
        Rt[:3, :3] = self.R; Rt[:3, 3] = self.T
        return Rt.transpose(0, 1)

    @property
    def full_proj_transform(self):
        return self.world_view_transform.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0)).squeeze(0)

    @property
    def camera_center(self):
        return self.world_view_transform.inverse()[3, :3]


def update_pose(cam):
    tau = torch.cat([cam.cam_trans_delta, cam.cam_rot_delta]).detach()
    T = torch.eye(4, device=tau.device); T[:3, :3] = cam.R; T[:3, 3] = cam.T
    new = SE3_exp(tau) @ T
    cam.R, cam.T = new[:3, :3], new[:3, 3]
    cam.cam_rot_delta.data.fill_(0); cam.cam_trans_delta.data.fill_(0)


class Gaussians:
    def __init__(self, sc, dev):
        t = lambda a: torch.tensor(a, device=dev)
        self.get_xyz, self.get_opacity, self.get_scaling = t(sc["means3D"]), t(sc["opacities"]), t(sc["scales"])
        self.get_rotation, self.get_features = t(sc["rotations"]), t(sc["shs"])
        self.active_sh_degree = 0


class Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False


def test_render_dict_and_custom_resolution():
    dev = "cuda"
    c = synth.make_camera("kitti", k=1)
    cam, pc = Cam(c, dev), Gaussians(synth.make_scene(30_000, c, seed=2), dev)
    bg = torch.zeros(3, device=dev)
    pkg = render(cam, pc, Pipe(), bg)
    assert set(pkg) == {"render", "viewspace_points", "visibility_filter", "radii", "depth", "opacity", "n_touched"}
    H, W = c.image_height, c.image_width
    assert pkg["render"].shape == (3, H, W) and pkg["depth"].shape == (1, H, W) and pkg["opacity"].shape == (1, H, W)
    assert pkg["radii"].dtype == torch.int32 and pkg["n_touched"].dtype == torch.int32
    assert torch.equal(pkg["visibility_filter"], pkg["radii"] > 0)
    assert 0 < int((pkg["n_touched"] > 0).sum()) <= int(pkg["visibility_filter"].sum())
    low = render_with_custom_resolution(cam, pc, Pipe(), bg, target_width=512, target_height=144)   # utils/init_pose.py:145
    assert low["depth"].shape == (1, 144, 512)
    # same FoV, lower resolution: the depth image is a down-sampled version of the full one (coarse check)
    full = torch.nn.functional.interpolate(pkg["depth"][None], size=(144, 512), mode="area")[0]
    m = (low["opacity"] > 0.9) & (torch.nn.functional.interpolate(pkg["opacity"][None], size=(144, 512), mode="area")[0] > 0.9)
    assert float(((low["depth"] - full).abs() / full.clamp_min(1.0))[m].median()) < 0.15
    # mask argument renders a subset
    sel = torch.zeros(30_000, dtype=torch.bool, device=dev); sel[::2] = True
    half = render(cam, pc, Pipe(), bg, mask=sel)
    assert half["radii"].shape == (15_000,)
    # gradients reach the screen-space points of the full render
    pkg["render"].sum().backward()
    assert pkg["viewspace_points"].grad is not None and float(pkg["viewspace_points"].grad.abs().sum()) > 0


def test_tracking_loop_recovers_a_perturbed_pose():
    """utils/slam_frontend.py:1468-1533 in miniature: Adam on (cam_rot_delta, cam_trans_delta), loss = L1 against the
    image rendered from the true pose, update_pose after every step."""
    dev = "cuda"
    c = synth.make_camera("mast3r_kitti")
    sc = synth.make_scene(40_000, c, seed=5)
    sc["opacities"] = np.clip(sc["opacities"] * 1.5, 0.3, 0.99).astype(np.float32)     # a solid scene to track against
    pc = Gaussians(sc, dev)
    bg = torch.zeros(3, device=dev)
    true_cam = Cam(c, dev)
    with torch.no_grad():
        target = render(true_cam, pc, Pipe(), bg)["render"].clone()
    cam = Cam(c, dev)
    tau0 = torch.tensor([0.03, -0.02, 0.04, math.radians(0.4), math.radians(-0.3), math.radians(0.2)], device=dev)
    T0 = SE3_exp(tau0) @ torch.eye(4, device=dev)
    cam.R, cam.T = T0[:3, :3].contiguous(), T0[:3, 3].contiguous()

    def pose_err():
        return float(torch.norm(cam.T - true_cam.T)), float(torch.norm(cam.R - true_cam.R))

    e0 = pose_err()
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": 0.003}, {"params": [cam.cam_trans_delta], "lr": 0.001}])
    losses = []
    for it in range(80):
        pkg = render(cam, pc, Pipe(), bg)
        loss = (pkg["opacity"] * (pkg["render"] - target).abs()).mean()      # get_loss_tracking_rgb, utils/slam_utils.py:53-62
        opt.zero_grad()
        loss.backward()
        opt.step()
        update_pose(cam)
        losses.append(float(loss))
    e1 = pose_err()
    assert losses[-1] < 0.35 * losses[0], (losses[0], losses[-1])
    assert e1[0] < 0.4 * e0[0] and e1[1] < 0.4 * e0[1], (e0, e1)
