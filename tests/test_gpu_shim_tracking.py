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


# The reference's own Camera / SE3_exp / update_pose (utils/camera_utils.py:8-166, utils/pose_utils.py:56-87) when
# /root/reference is mounted, else the restatements that tests/test_reference_pin.py pins against them.
import ref_conventions as rc

_PU, _SU, _CU, REF_SOURCE = rc.load()
SE3_exp, update_pose = _PU.SE3_exp, _PU.update_pose
Gaussians, Pipe = rc.Gaussians, rc.Pipe


def Cam(c: synth.Cam, dev):
    return rc.make_camera(_CU, c, dev)


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


def test_custom_resolution_render_matches_the_oracle_at_that_resolution():
    """utils/init_pose.py:145 renders a 512x144 depth image with the full-resolution camera's FoV and matrices; the same
    call through the oracle (W, H overridden, everything else the camera's) must agree: radii bit-exact, images to 1e-5 on
    the pixels that are not within float noise of a blend decision."""
    import oracle
    dev = "cuda"
    c = synth.make_camera("kitti", k=2)
    sc = synth.make_scene(40_000, c, seed=21)
    cam, pc = Cam(c, dev), Gaussians(sc, dev)
    bg = torch.zeros(3, device=dev)
    with torch.no_grad():
        low = render_with_custom_resolution(cam, pc, Pipe(), bg, target_width=512, target_height=144)
    fwd = oracle.rasterize_forward(sc["means3D"], sc["opacities"], sc["scales"], sc["rotations"], sc["shs"], None, None,
                                   viewmatrix=cam.world_view_transform.cpu().numpy(), projmatrix=cam.full_proj_transform.cpu().numpy(),
                                   campos=cam.camera_center.cpu().numpy(),      # the very matrices the shim hands to the rasterizer
                                   bg=np.zeros(3, np.float32), W=512, H=144, tanfovx=c.tanfovx, tanfovy=c.tanfovy, sh_degree=0)
    np.testing.assert_array_equal(low["radii"].cpu().numpy(), fwd["radii"])
    ok = fwd["margin"] > 1e-5
    assert ok.mean() > 0.995
    for key, ref in (("render", fwd["color"]), ("depth", fwd["depth"]), ("opacity", fwd["opacity"])):
        got = low[key].cpu().numpy()
        assert np.abs(got[:, ok] - ref[:, ok]).max() <= 1e-5 * max(1.0, float(np.abs(ref).max())), key
    touched_ok = fwd["n_touched"] == low["n_touched"].cpu().numpy()
    assert touched_ok.mean() > 0.999            # n_touched differs only for Gaussians that reach a knife-edge pixel


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
    cam.update_RT(T0[:3, :3].contiguous(), T0[:3, 3].contiguous())

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
        with torch.no_grad():
            opt.step()
            update_pose(cam)
        losses.append(float(loss))
    e1 = pose_err()
    assert losses[-1] < 0.35 * losses[0], (losses[0], losses[-1])
    assert e1[0] < 0.4 * e0[0] and e1[1] < 0.4 * e0[1], (e0, e1)
