"""The device-resident tracking loop (lvdgs.tracking.PoseTracker: rasterizer + fused loss + lvdgs_pose_step) against the
reference's formulation of the same loop -- torch.optim.Adam on (cam_rot_delta, cam_trans_delta, exposure_a,
exposure_b) + update_pose (utils/slam_frontend.py:1466-1521, utils/pose_utils.py:56-87): the reference's own Camera /
SE3_exp / update_pose when /root/reference is mounted, else the restatements pinned against them
(tests/ref_conventions.py, tests/test_reference_pin.py)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from lvdgs import synth, _native
from lvdgs import tracking as trk
from gaussian_splatting.gaussian_renderer import render
from test_gpu_shim_tracking import Cam, Gaussians, Pipe, SE3_exp, update_pose

pytestmark = pytest.mark.gpu


def test_pose_step_matches_adam_plus_update_pose():
    dev = "cuda"
    c = synth.make_camera("kitti", k=3)
    cam = Cam(c, dev)
    cam.exposure_a = torch.nn.Parameter(torch.tensor([0.05], device=dev))
    cam.exposure_b = torch.nn.Parameter(torch.tensor([-0.01], device=dev))
    lrs = (0.003, 0.001, 0.01)
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": lrs[0]}, {"params": [cam.cam_trans_delta], "lr": lrs[1]},
                            {"params": [cam.exposure_a], "lr": lrs[2]}, {"params": [cam.exposure_b], "lr": lrs[2]}])
    L = _native.lib()
    tr = trk.PoseTracker(16, c.image_width, c.image_height, c.tanfovx, c.tanfovy, device=dev)
    tr.set_camera(cam.R, cam.T, cam.projection_matrix, 0.05, -0.01)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cpu").manual_seed(4)
    for step in range(1, 6):
        g_tau = (torch.randn(6, generator=g) * (10.0 if step < 5 else 1e-9)).to(dev)      # last step: |tau| ~ 0 (series branch)
        g_exp = torch.randn(2, generator=g).to(dev)
        cam.cam_trans_delta.grad = g_tau[:3].clone(); cam.cam_rot_delta.grad = g_tau[3:].clone()
        cam.exposure_a.grad = g_exp[:1].clone(); cam.exposure_b.grad = g_exp[1:].clone()
        opt.step()
        tau_norm = float(torch.cat([cam.cam_trans_delta, cam.cam_rot_delta]).norm())
        with torch.no_grad():
            update_pose(cam)
        _native.check(L.lvdgs_pose_step(_native.ptr(tr.state), _native.ptr(g_tau), _native.ptr(g_exp), *lrs, 0.9, 0.999, 1e-8,
                                        step, 1e-4, stream), "lvdgs_pose_step")
        st = tr.state.cpu().numpy()
        np.testing.assert_allclose(st[trk._R:trk._R + 9].reshape(3, 3), cam.R.cpu().numpy(), atol=2e-6)
        np.testing.assert_allclose(st[trk._T:trk._T + 3], cam.T.cpu().numpy(), atol=2e-6)
        np.testing.assert_allclose(st[trk._VIEW:trk._VIEW + 16].reshape(4, 4), cam.world_view_transform.cpu().numpy(), atol=2e-6)
        np.testing.assert_allclose(st[trk._PROJ:trk._PROJ + 16].reshape(4, 4), cam.full_proj_transform.cpu().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(st[trk._CAMPOS:trk._CAMPOS + 3], cam.camera_center.cpu().numpy(), atol=1e-5)
        np.testing.assert_allclose(st[trk._EXPO:trk._EXPO + 2], [float(cam.exposure_a), float(cam.exposure_b)], atol=1e-6)
        assert abs(float(st[trk._TAUN]) - tau_norm) < 1e-6
        assert int(tr.state.view(torch.int32)[trk._CONV]) == int(tau_norm < 1e-4)


def test_native_tracking_loop_follows_the_reference_loop():
    """Same scene, same start pose, same learning rates: the device-resident loop and the plugin + torch loop must walk the
    same trajectory (both are float32; they differ only by summation order), and both must recover the pose."""
    dev = "cuda"
    c = synth.make_camera("mast3r_kitti")
    sc = synth.make_scene(40_000, c, seed=5)
    sc["opacities"] = np.clip(sc["opacities"] * 1.5, 0.3, 0.99).astype(np.float32)
    pc = Gaussians(sc, dev)
    bg = torch.zeros(3, device=dev)
    true_cam = Cam(c, dev)
    with torch.no_grad():
        target = render(true_cam, pc, Pipe(), bg)["render"].clone()
    tau0 = torch.tensor([0.03, -0.02, 0.04, math.radians(0.4), math.radians(-0.3), math.radians(0.2)], device=dev)
    T0 = SE3_exp(tau0)
    iters = 40
    # reference formulation
    cam = Cam(c, dev)
    cam.update_RT(T0[:3, :3].contiguous(), T0[:3, 3].contiguous())
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": 0.003}, {"params": [cam.cam_trans_delta], "lr": 0.001}])
    ref_losses = []
    for it in range(iters):
        pkg = render(cam, pc, Pipe(), bg)
        loss = (pkg["opacity"] * (pkg["render"] - target).abs()).mean()
        opt.zero_grad(); loss.backward()
        with torch.no_grad():
            opt.step(); update_pose(cam)
        ref_losses.append(float(loss))
    # device-resident loop
    tr = trk.PoseTracker(40_000, c.image_width, c.image_height, c.tanfovx, c.tanfovy, device=dev, lr_rot=0.003, lr_trans=0.001,
                         optimise_exposure=False, rgb_boundary_threshold=-1.0)
    args = (pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features, target)
    # (a) one iteration from the start pose: same loss, same pose gradient as the plugin + autograd path
    cam1 = Cam(c, dev)
    cam1.update_RT(T0[:3, :3].contiguous(), T0[:3, 3].contiguous())
    pkg = render(cam1, pc, Pipe(), bg)
    loss1 = (pkg["opacity"] * (pkg["render"] - target).abs()).mean()
    loss1.backward()
    tr.set_camera(T0[:3, :3], T0[:3, 3], true_cam.projection_matrix)
    out1 = tr.track(*args, iters=1, stop_when_converged=False)
    assert abs(float(out1["loss"]) - float(loss1)) < 1e-5 * float(loss1)
    g_nat = tr.eng.slots[0].g_tau.cpu().numpy()
    g_ref = np.concatenate([cam1.cam_trans_delta.grad.cpu().numpy(), cam1.cam_rot_delta.grad.cpu().numpy()])
    np.testing.assert_allclose(g_nat, g_ref, rtol=2e-4, atol=1e-6 * np.abs(g_ref).max())
    # (b) the whole loop.  Adam's first steps are lr * g / |g|, so float noise in a near-zero gradient component moves a
    # step by up to 2 lr: the two trajectories agree to a few learning rates, not to float precision
    tr.set_camera(T0[:3, :3], T0[:3, 3], true_cam.projection_matrix)
    out = tr.track(*args, iters=iters, stop_when_converged=False)
    assert out["steps"] == iters
    e_ref = float(torch.norm(cam.T - true_cam.T)), float(torch.norm(cam.R - true_cam.R))
    e_nat = float(torch.norm(tr.T - true_cam.T)), float(torch.norm(tr.R - true_cam.R))
    e_0 = float(torch.norm(T0[:3, 3] - true_cam.T)), float(torch.norm(T0[:3, :3] - true_cam.R))
    assert e_nat[0] < 0.4 * e_0[0] and e_nat[1] < e_0[1], (e_0, e_nat, e_ref)          # 40 iterations: translation mostly recovered
    assert float(torch.norm(tr.T - cam.T)) < 0.05 * e_0[0] and float(torch.norm(tr.R - cam.R)) < 0.05 * e_0[1], (e_0, e_ref, e_nat)
    assert float(out["loss"]) < 0.35 * ref_losses[0] and abs(float(out["loss"]) - ref_losses[-1]) < 0.25 * ref_losses[-1]


def test_native_tracking_stops_on_convergence():
    dev = "cuda"
    c = synth.make_camera("mast3r_kitti")
    sc = synth.make_scene(20_000, c, seed=6)
    pc = Gaussians(sc, dev)
    true_cam = Cam(c, dev)
    with torch.no_grad():
        target = render(true_cam, pc, Pipe(), torch.zeros(3, device=dev))["render"].clone()
    # already at the optimum, huge threshold: the first step reports convergence, the loop stops after the next forward
    tr = trk.PoseTracker(20_000, c.image_width, c.image_height, c.tanfovx, c.tanfovy, device=dev, converged_threshold=1.0,
                         optimise_exposure=True, rgb_boundary_threshold=-1.0)
    tr.set_camera(true_cam.R, true_cam.T, true_cam.projection_matrix)
    out = tr.track(pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features, target, iters=50)
    assert out["steps"] == 1


def test_native_rgbd_tracking_iteration_matches_autograd_through_the_reference_loss():
    """RGB-D configs: one PoseTracker iteration with gt_depth = get_loss_tracking_rgbd (utils/slam_utils.py:65-83; the
    reference's function when /root/reference is mounted, else the restatement pinned against it) + autograd through the
    plugin: same loss, same pose gradient (the depth gradient enters the pose-only backward)."""
    import ref_conventions as rc
    dev = "cuda"
    c = synth.make_camera("mast3r_kitti")
    sc = synth.make_scene(30_000, c, seed=8)
    sc["opacities"] = np.clip(sc["opacities"] * 1.6, 0.4, 0.99).astype(np.float32)
    pc = Gaussians(sc, dev)
    bg = torch.zeros(3, device=dev)
    true_cam = Cam(c, dev)
    with torch.no_grad():
        pkg0 = render(true_cam, pc, Pipe(), bg)
        target, target_depth = pkg0["render"].clone(), pkg0["depth"].clone()
    rng = np.random.default_rng(3)
    target_depth = target_depth * torch.tensor(rng.uniform(0.9, 1.1, tuple(target_depth.shape)).astype(np.float32), device=dev)
    target_depth[0, :7, :] = 0.0                                   # pixels without a depth measurement
    tau0 = torch.tensor([0.02, -0.03, 0.03, math.radians(0.3), math.radians(-0.2), math.radians(0.25)], device=dev)
    T0 = SE3_exp(tau0)
    cam1 = Cam(c, dev)
    cam1.update_RT(T0[:3, :3].contiguous(), T0[:3, 3].contiguous())
    cam1.original_image = target
    cam1.mono_depth = target_depth[0].cpu().numpy()
    cam1.grad_mask = torch.ones(1, c.image_height, c.image_width, device=dev)
    config = {"Training": {"monocular": False, "rgb_boundary_threshold": 0.01, "alpha": 0.9}, "Dataset": {"depth_loss": True}}
    pkg = render(cam1, pc, Pipe(), bg)
    _, su, _, _ = rc.load()
    loss1 = su.get_loss_tracking_rgbd(config, pkg["render"], pkg["depth"], pkg["opacity"], cam1)
    loss1.backward()
    tr = trk.PoseTracker(30_000, c.image_width, c.image_height, c.tanfovx, c.tanfovy, device=dev, optimise_exposure=False,
                         rgb_boundary_threshold=0.01)
    tr.set_camera(T0[:3, :3], T0[:3, 3], true_cam.projection_matrix)
    out = tr.track(pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features, target, iters=1,
                   stop_when_converged=False, gt_depth=target_depth, alpha=0.9)
    loss_rgbd = float(out["loss"])                 # (out["loss"] is a view of the tracker's result buffer: read it now)
    assert abs(loss_rgbd - float(loss1)) < 2e-5 * float(loss1)
    g_nat = tr.eng.slots[0].g_tau.cpu().numpy()
    g_ref = np.concatenate([cam1.cam_trans_delta.grad.cpu().numpy(), cam1.cam_rot_delta.grad.cpu().numpy()])
    np.testing.assert_allclose(g_nat, g_ref, rtol=5e-4, atol=2e-6 * np.abs(g_ref).max())
    # and the rgb-only iteration on the same frame differs (the depth term is live)
    tr.set_camera(T0[:3, :3], T0[:3, 3], true_cam.projection_matrix)
    out_rgb = tr.track(pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features, target, iters=1,
                       stop_when_converged=False)
    assert abs(float(out_rgb["loss"]) - loss_rgbd) > 1e-3 * loss_rgbd
